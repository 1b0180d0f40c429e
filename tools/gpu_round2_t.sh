#!/bin/bash
# 1-GPU visit: in-library streamed merge (parity + timing), mixed CS / rebuilt-string batch
timeout 900 python -m pytest tests/test_gpu_merge.py::test_streamed_merge_of_runs_that_do_not_fit_together tests/test_gpu_string_codecs.py::test_batch_with_cs_coded_streams_and_rebuilt_pax_strings -q -x 2>&1 | tail -12
timeout 600 python bench.py --workload compaction --compaction-window 400000 --stream-ranges 6 --verify --steps 2 --warmup 1 2>gpurun_out/t_small.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streamed (C) verify', d.get('parity'), d['ms_per_step'])"
tail -3 gpurun_out/t_small.err
for r in 8 24; do
timeout 900 python bench.py --workload compaction --compaction-window 24000000 --stream-ranges $r --steps 3 --warmup 1 2>gpurun_out/t_$r.err | tee gpurun_out/compaction_streamed_c_$r.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streamed (C) ranges $r', d['config']['input_rows'], d['ms_per_step'], d['value'])"
done
tail -2 gpurun_out/t_8.err
