#!/bin/bash
# N-GPU visit (gpurun --gpus N): range-partitioned compaction merge (parity at small size, then scale), the scan
# bench line and config 3 at N ranks.
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/bench_compaction.py --runs 8 --window 400000 --verify 2>gpurun_out/mg_cmp_small.err | tee gpurun_out/compaction_n${N}_small.json | cut -c1-1400
tail -2 gpurun_out/mg_cmp_small.err
$TR --master-port 29512 tools/bench_compaction.py --runs 8 --window 4000000 2>gpurun_out/mg_cmp.err | tee gpurun_out/compaction_n${N}.json | cut -c1-1400
tail -2 gpurun_out/mg_cmp.err
$TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/mg_bench.err | tee gpurun_out/bench_n${N}.json | cut -c1-700
tail -2 gpurun_out/mg_bench.err
$TR --master-port 29514 tools/bench_config.py --config 3 --rows 8000000 --tile 4 --cpu-blocks 0 2>gpurun_out/mg_cfg3.err | tee gpurun_out/cfg3_n${N}.json | cut -c1-700
tail -2 gpurun_out/mg_cfg3.err
