#!/bin/bash
# One GPU-box visit: bench lines (ours + reference arm), ncu launch list of the bench command and
# one --set full capture of each scan kernel. Outputs land in gpurun_out/ with the given tag.
tag=${1:-r1_v7}
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -c 1500 gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err
cut -c1-400 gpurun_out/bench_${tag}_ref.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none -k regex:obgpu -c 40 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
for k in project count; do
  ncu --set full --clock-control none --import-source on -k regex:obgpu_${k}_kernel -s 3 -c 1 \
      -o gpurun_out/prof_${tag}_${k} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 \
      > gpurun_out/ncu_full_${k}.log 2>&1
  tail -1 gpurun_out/ncu_full_${k}.log
done
