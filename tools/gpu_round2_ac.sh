#!/bin/bash
# whole GPU suite after the hash-slot screen of string equality leaves + cfg3 (133-row blocks) A/B timing and per-kernel times
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/bench_config.py --rows 4000000 --tile 4 --rows-per-block 133 --cpu-blocks 0 2>gpurun_out/ac.err | tee gpurun_out/cfg3_ac.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('cfg3/133', d['ms_per_step'], d['value'], d['roofline'].get('frac'), d['roofline'].get('kernel_ms'))"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:obgpu -s 8 -c 4 --csv --log-file gpurun_out/launches_cfg3_ac.csv python tools/bench_config.py --rows 4000000 --tile 4 --rows-per-block 133 --steps 2 --cpu-blocks 0 > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_cfg3_ac.csv | cut -d, -f5,13- | tail -8
