#!/bin/bash
# parity (both dispatches) + the new datum / pipeline / adapter paths + the full bench line
tag=${1:-f}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
OBGPU_PIPE=1 timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full_${tag}.json 2> gpurun_out/bench_full_${tag}.err
tail -c 6000 gpurun_out/bench_full_${tag}.json | cut -c1-6000; tail -5 gpurun_out/bench_full_${tag}.err
