#!/bin/bash
# cfg4 (TPC-H Q6 columns as a CS column group): RAW streams vs the codecs the encoder detection picks (decoded at batch open)
for m in raw detect; do
timeout 600 python tools/bench_config.py --config 4 --rows 25000000 --tile 4 --cs-streams $m --steps 10 --warmup 3 2>gpurun_out/v_$m.err | tee gpurun_out/cfg4_$m.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('cfg4 $m', d['ms_per_step'], d['value'], d['roofline']['frac'], 'B/row', d['config']['bytes_per_row_in'], 'open_ms', d['open_ms'], 'q6', d.get('q6_revenue_x10000'))"
tail -2 gpurun_out/v_$m.err
done
