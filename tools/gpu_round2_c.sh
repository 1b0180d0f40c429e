#!/bin/bash
# round-2 iteration visit: parity under both dispatches, launch list at 125 M rows, source-level ncu of the pipe kernels
tag=${1:-c}
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
OBGPU_PIPE=1 timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:obgpu -s 6 -c 4 --csv --log-file gpurun_out/launches_cfg3_${tag}.csv python bench.py --rows 125000000 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1 --e2e-one-tile > /dev/null 2>gpurun_out/ncu_${tag}.err
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
for k in ${NCU_KERNELS:-obgpu_count_pipe_kernel obgpu_project_pipe_kernel}; do
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -f -o gpurun_out/${k}_${tag} python bench.py --rows 31250000 --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1 --e2e-one-tile > /dev/null 2>>gpurun_out/ncu_${tag}.err
done
python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1 --e2e-one-tile 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 1B', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
