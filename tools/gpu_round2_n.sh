#!/bin/bash
# 1-GPU visit: string codecs (all cases), adapter, cfg3 timing after the string screen, whole gpu suite
timeout 900 python -m pytest tests/test_gpu_string_codecs.py -q 2>&1 | tail -30
timeout 300 tests/cpp/test_host_adapter | tail -6
timeout 900 python bench.py --rows 250000000 --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1 --e2e-one-tile 2>gpurun_out/n_bench.err | tee gpurun_out/bench_cfg3_n_250m.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('cfg3 250M rows', d['ms_per_step'], d['value'], d['roofline']['frac'])"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
