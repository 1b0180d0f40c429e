#!/bin/bash
# 2-GPU visit: range-partitioned compaction with phase B on every rank -- parity (merged stream, column checksums summed over the ranks)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 tools/bench_compaction.py --runs 8 --window 400000 --verify > gpurun_out/af_small.out 2>gpurun_out/af_small.err
grep '^{' gpurun_out/af_small.out | tail -1 > gpurun_out/r2_compaction_phaseb_n2_verify.json
python -c "
import json
d=json.load(open('gpurun_out/r2_compaction_phaseb_n2_verify.json')); print('n2 verify', d['parity'], d['phase_b']['column_checksums_match'], d['phase_b']['encode_ms'], d['phase_b']['n_blocks'], d['ms_per_step'])"
tail -2 gpurun_out/af_small.err
