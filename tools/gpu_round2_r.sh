#!/bin/bash
# N-GPU visit (N = 8): compaction with the in-library exchange at the larger size, cfg3 strong scaling line. Bounded: every command has its own timeout.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29531 bench.py --workload compaction --gpus $N --compaction-window 24000000 --steps 5 --warmup 2 2>gpurun_out/r_cmp.err | tee gpurun_out/compaction_n${N}_w24m.json | cut -c1-1400
tail -3 gpurun_out/r_cmp.err
timeout 300 $TR --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 --no-secondary --e2e-steps 1 2>gpurun_out/r_bench.err | tee gpurun_out/bench_cfg3_n${N}.json | cut -c1-2600
tail -3 gpurun_out/r_bench.err
