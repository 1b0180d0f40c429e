#!/bin/bash
# e2e (host buffers in, host vectors out): NUMA binding on/off, worker streams, pipeline depth
for nb in "" "--no-numa-bind"; do for w in 3 5; do for b in 12 24; do
  python bench.py --no-cpu-baseline --steps 3 --e2e-steps 3 --e2e-workers $w --e2e-batches $b $nb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('numa', d['config']['host_numa_node'], 'workers', $w, 'batches', e['page_batches'], 'e2e ms', round(e['ms_per_step'],2), 'Grows/s', round(e['value']/1e9,3), 'kernel ms', round(d['roofline']['kernel_ms'],4))"
done; done; done
