#!/bin/bash
# 1-GPU visit: compaction v3 parity + timing + source-level ncu of the two merge kernels
timeout 600 python -m pytest tests/test_gpu_merge.py -x -q 2>&1 | tail -3
timeout 300 tests/cpp/test_partition_merger | tail -1
timeout 600 python tools/bench_compaction.py --runs 8 --window 400000 --verify 2>gpurun_out/l_small.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('verify', d['parity'], d['ms_per_step'])"
for w in 4000000 24000000; do
  timeout 900 python bench.py --workload compaction --compaction-window $w --steps 5 --warmup 2 2>gpurun_out/l_$w.err | tee gpurun_out/compaction_l_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('v3', d['config']['input_rows'], d['ms_per_step'], d['phases_ms']['decode_runs'], d['phases_ms']['exchange_plus_merge'], d['roofline']['frac'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'bucket|fuse|head|sample' -c 5 --csv --log-file gpurun_out/launches_compaction_l.csv python bench.py --workload compaction --compaction-window 24000000 --steps 1 --warmup 0 > /dev/null 2>gpurun_out/ncu_l.err
for k in bucket_merge_kernel fuse_bucket_kernel; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/${k}_l python bench.py --workload compaction --compaction-window 4000000 --steps 1 --warmup 0 > /dev/null 2>>gpurun_out/ncu_l.err
done
tail -2 gpurun_out/ncu_l.err
