#!/bin/bash
# per-kernel timeline of one compaction step (all kernels) at two sizes
for w in 4000000 24000000; do
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_compaction_j_$w.csv python bench.py --workload compaction --compaction-window $w --steps 1 --warmup 1 > /dev/null 2>gpurun_out/ncu_j.err
tail -2 gpurun_out/ncu_j.err
done
