#!/bin/bash
# 1-GPU visit: zero-copy host scan (parity + e2e both modes on cfg3, smaller row count for the device-resident part), adapter iterator
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q 2>&1 | tail -5
timeout 300 tests/cpp/test_host_adapter | tail -3
timeout 1200 python bench.py --rows 250000000 --steps 3 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 2 2>gpurun_out/o_bench.err | tee gpurun_out/bench_cfg3_o.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('device', d['ms_per_step'], d['value'])
for k in ('e2e','e2e_staged','e2e_zero_copy'):
    if k in d: print(k, d[k]['value'], d[k]['ms_per_step'], d[k]['h2d_bytes_per_step'], d[k]['d2h_bytes_per_step'], d[k].get('mode','')[:40])"
tail -3 gpurun_out/o_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
