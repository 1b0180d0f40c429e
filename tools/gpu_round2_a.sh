#!/bin/bash
# round-2 GPU visit A: parity tests + cfg3 headline bench with the round-1 kernels (baseline for the new ones)
tag=${1:-a}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-secondary --e2e-steps 1 > gpurun_out/bench_cfg3_${tag}.json 2> gpurun_out/bench_cfg3_${tag}.err
tail -c 3000 gpurun_out/bench_cfg3_${tag}.json; tail -5 gpurun_out/bench_cfg3_${tag}.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:obgpu -c 12 --csv --log-file gpurun_out/launches_cfg3_${tag}.csv python bench.py --rows 125000000 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1 --e2e-one-tile > /dev/null 2>gpurun_out/ncu_${tag}.err
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
