#!/bin/bash
# 1-GPU visit: bucket merge v2 (rank by advance, heads in the bucket kernel, per-bucket fuse): parity, timing, per-kernel timeline
timeout 900 python -m pytest tests/test_gpu_merge.py tests/test_gpu_block_api.py -x -q 2>&1 | tail -5
timeout 300 tests/cpp/test_partition_merger | tail -2
timeout 600 python tools/bench_compaction.py --runs 8 --window 400000 --verify 2>gpurun_out/k_small.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('verify', d['parity'], d['ms_per_step'])"
for w in 4000000 24000000; do
  timeout 900 python bench.py --workload compaction --compaction-window $w --steps 5 --warmup 2 2>gpurun_out/k_bucket_$w.err | tee gpurun_out/compaction_k_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('bucket v2', d['config']['input_rows'], d['ms_per_step'], d['phases_ms']['decode_runs'], d['phases_ms']['exchange_plus_merge'], d['roofline']['frac'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'bucket|fuse|head|sample' -c 12 --csv --log-file gpurun_out/launches_compaction_k.csv python bench.py --workload compaction --compaction-window 24000000 --steps 1 --warmup 1 > /dev/null 2>gpurun_out/ncu_k.err
tail -2 gpurun_out/ncu_k.err
