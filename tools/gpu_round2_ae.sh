#!/bin/bash
# final 1-GPU visit: the bench line of the final code, the reference arm, the launch list of a reduced-size run of the same command
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -c 2500 gpurun_out/r2_bench_final.json
tail -3 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_reference_arm.json 2>/dev/null
cut -c1-500 gpurun_out/r2_bench_final_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:obgpu -c 24 --csv \
  --log-file gpurun_out/r2_launches_cfg3_final.csv python bench.py --rows 125000000 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 0 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_launches_cfg3_final.csv | cut -d, -f5,13- | grep "duration\|dram" | tail -12
