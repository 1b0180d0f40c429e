#!/bin/bash
# e2e (host buffers in, host vectors out) as a function of pipeline depth and worker streams
for w in 3 5; do for b in 12 24 40; do
  python bench.py --no-cpu-baseline --steps 3 --e2e-steps 3 --e2e-workers $w --e2e-batches $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('workers', $w, 'batches', e['page_batches'], 'e2e ms', round(e['ms_per_step'],2), 'Grows/s', round(e['value']/1e9,3), 'kernel ms', round(d['roofline']['kernel_ms'],4))"
done; done
