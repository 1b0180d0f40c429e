#!/bin/bash
# zero-copy e2e sweep: worker streams / page batches
for wk in 3 6 10; do for nb in 14 28; do
timeout 600 python bench.py --rows 125000000 --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 2 --e2e-workers $wk --e2e-batches $nb 2>gpurun_out/p_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
z = d['e2e'] if 'zero' in d['e2e'].get('mode','') else d.get('e2e_zero_copy')
s = d.get('e2e_staged', d['e2e'])
print('workers $wk batches $nb zero_copy', round(z['value']/1e9,3), 'ms', round(z['ms_per_step'],1), ' staged', round(s['value']/1e9,3))"
done; done
