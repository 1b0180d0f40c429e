#!/bin/bash
# 1-GPU visit: new merge shape tests, mixed CS + rebuilt-string batch, streamed merge timing with pinned fetch
timeout 900 python -m pytest tests/test_gpu_merge.py tests/test_gpu_string_codecs.py -q -x 2>&1 | tail -12
timeout 900 python bench.py --workload compaction --compaction-window 24000000 --stream-ranges 8 --steps 3 --warmup 1 2>gpurun_out/s_8.err | tee gpurun_out/compaction_streamed_pinned_8.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('streamed ranges 8 (pinned fetch)', d['config']['input_rows'], d['ms_per_step'], d['value'])"
tail -2 gpurun_out/s_8.err
