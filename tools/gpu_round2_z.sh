#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_device_encoder.py -q -x 2>&1 | tail -5
timeout 300 python tools/bench_encode.py 2>&1 | tail -1 | tee gpurun_out/r2_encode_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:obgpu_encode -s 3 -c 1 --csv --log-file gpurun_out/r2_encode_launches.csv python tools/bench_encode.py --steps 2 --warmup 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_encode_launches.csv | cut -d, -f13- | tail -5


