#!/bin/bash
# N-GPU visit: compaction with the in-library NCCL exchange (parity at small size, then timing) + the cfg3 strong-scaling line
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_merge.py tests/test_gpu_block_api.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -4
timeout 300 tests/cpp/test_host_adapter | tail -2
$TR --master-port 29511 tools/bench_compaction.py --runs 8 --window 400000 --verify 2>gpurun_out/mg_cmp_small.err | tee gpurun_out/compaction_n${N}_small.json | cut -c1-1800
tail -3 gpurun_out/mg_cmp_small.err
$TR --master-port 29512 bench.py --workload compaction --gpus $N --steps 5 --warmup 2 2>gpurun_out/mg_cmp.err | tee gpurun_out/compaction_n${N}.json | cut -c1-1800
tail -3 gpurun_out/mg_cmp.err
python bench.py --workload compaction --steps 5 --warmup 2 2>gpurun_out/mg_cmp1.err | tee gpurun_out/compaction_n1.json | cut -c1-1500
$TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-secondary --e2e-one-tile 2>gpurun_out/mg_bench.err | tee gpurun_out/bench_cfg3_n${N}.json | cut -c1-2500
tail -3 gpurun_out/mg_bench.err
