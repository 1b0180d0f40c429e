#!/bin/bash
# 1-GPU visit: single-pass K-way merge vs the pairwise passes (parity + timing at two sizes), datum test
timeout 900 python -m pytest tests/test_gpu_merge.py tests/test_gpu_block_api.py -x -q 2>&1 | tail -5
timeout 300 tests/cpp/test_partition_merger | tail -2
timeout 600 python tools/bench_compaction.py --runs 8 --window 400000 --verify 2>gpurun_out/i_small.err | cut -c1-1600
for w in 4000000 24000000; do
  OBGPU_MERGE_PAIRWISE=1 timeout 900 python bench.py --workload compaction --compaction-window $w --steps 5 --warmup 2 2>gpurun_out/i_pair_$w.err | tee gpurun_out/compaction_pairwise_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('pairwise', d['config']['input_rows'], d['ms_per_step'], d['phases_ms']['decode_runs'], d['phases_ms']['exchange_plus_merge'], d['roofline']['frac'])"
  timeout 900 python bench.py --workload compaction --compaction-window $w --steps 5 --warmup 2 2>gpurun_out/i_bucket_$w.err | tee gpurun_out/compaction_bucket_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('bucket  ', d['config']['input_rows'], d['ms_per_step'], d['phases_ms']['decode_runs'], d['phases_ms']['exchange_plus_merge'], d['roofline']['frac'])"
done
tail -3 gpurun_out/i_*.err | tail -20
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'mrg|prefix|cub|decode' -c 60 --csv --log-file gpurun_out/launches_compaction_i.csv python bench.py --workload compaction --compaction-window 4000000 --steps 1 --warmup 1 > /dev/null 2>gpurun_out/ncu_i.err
tail -2 gpurun_out/ncu_i.err
