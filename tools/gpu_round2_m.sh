#!/bin/bash
# 1-GPU visit: string codecs on the device, compaction timing after the per-thread share change, whole gpu suite
timeout 900 python -m pytest tests/test_gpu_string_codecs.py -x -q 2>&1 | tail -30
for w in 4000000 24000000; do
  timeout 900 python bench.py --workload compaction --compaction-window $w --steps 5 --warmup 2 2>gpurun_out/m_$w.err | tee gpurun_out/compaction_m_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('v4', d['config']['input_rows'], d['ms_per_step'], d['phases_ms']['decode_runs'], d['phases_ms']['exchange_plus_merge'], d['roofline']['frac'])"
done
timeout 600 python tools/bench_compaction.py --runs 8 --window 400000 --verify 2>gpurun_out/m_small.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('verify', d['parity'], d['ms_per_step'])"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
