#!/bin/bash
# GPU-box visit used while iterating: parity tests, cfg3 and cfg2 device-resident numbers, per-kernel times.
tag=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/bench_config.py --rows 8000000 --tile 4 --cpu-blocks 0 2>gpurun_out/cfg3.err | tee gpurun_out/cfg3_${tag}.json | cut -c1-900
python bench.py --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'])"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:obgpu -s 5 -c 4 --csv --log-file gpurun_out/launches_cfg3_${tag}.csv python tools/bench_config.py --rows 8000000 --tile 4 --steps 2 --cpu-blocks 0 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_cfg3_${tag}.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<len(H): continue
    d=dict(zip(H,r))
    agg.setdefault((d['ID'], d['Kernel Name'][:28], d['Grid Size']),{})[d['Metric Name'].split('__')[1][:14]]=d['Metric Value']
for k,v in agg.items(): print(k, v)
PY
