#!/bin/bash
# 1-GPU visit: streamed merge (parity test, verify at small size, timing at the larger size), per-kernel timeline of the compaction
timeout 600 python -m pytest tests/test_gpu_merge.py -x -q 2>&1 | tail -6
timeout 600 python bench.py --workload compaction --compaction-window 400000 --stream-ranges 6 --verify --steps 2 --warmup 1 2>gpurun_out/q_small.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streamed verify', d.get('parity'), d['ms_per_step'])"
tail -3 gpurun_out/q_small.err
for r in 8 16; do
timeout 900 python bench.py --workload compaction --compaction-window 24000000 --stream-ranges $r --steps 3 --warmup 1 2>gpurun_out/q_$r.err | tee gpurun_out/compaction_streamed_$r.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streamed ranges $r', d['config']['input_rows'], d['ms_per_step'], d['value'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'bucket|fuse|sample|decode' -c 16 --csv --log-file gpurun_out/launches_compaction_q.csv python bench.py --workload compaction --compaction-window 24000000 --steps 1 --warmup 0 > /dev/null 2>gpurun_out/ncu_q.err
tail -2 gpurun_out/ncu_q.err
