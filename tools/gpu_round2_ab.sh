#!/bin/bash
# cfg3: per-source-line instruction counts of the two pipelined kernels (ncu --set full with source), plus the plain timing
timeout 300 python tools/bench_config.py --rows 4000000 --tile 4 --rows-per-block 133 --cpu-blocks 0 2>gpurun_out/ab.err | tee gpurun_out/cfg3_ab.json | cut -c1-600
for k in count_pipe project_pipe; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:obgpu_${k} -s 3 -c 1 -o gpurun_out/r2_cfg3_${k}_full python tools/bench_config.py --rows 4000000 --tile 4 --rows-per-block 133 --steps 2 --cpu-blocks 0 > gpurun_out/ncu_ab_${k}.log 2>&1
  tail -1 gpurun_out/ncu_ab_${k}.log
done
