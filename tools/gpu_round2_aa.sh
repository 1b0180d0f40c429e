#!/bin/bash
# 1-GPU visit: C++ adapter tests (column groups), compaction with phase B on the device (verify size + the 116 M-row size)
timeout 600 python -m pytest tests/test_host_adapter.py tests/test_gpu_device_encoder.py -x -q 2>&1 | tail -12
timeout 600 python tools/bench_compaction.py --runs 8 --window 400000 --verify 2>gpurun_out/aa_small.err | tee gpurun_out/r2_compaction_phaseb_verify.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('verify', d['parity'], d['phase_b'])"
tail -3 gpurun_out/aa_small.err
for w in 24000000; do
  timeout 900 python bench.py --workload compaction --compaction-window $w --steps 5 --warmup 2 2>gpurun_out/aa_$w.err | tee gpurun_out/r2_compaction_phaseb_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('cfg5', d['config']['input_rows'], d['ms_per_step'], d['phases_ms']['decode_runs'], d['phases_ms']['exchange_plus_merge'], d['phase_b'])"
  tail -3 gpurun_out/aa_$w.err
done
