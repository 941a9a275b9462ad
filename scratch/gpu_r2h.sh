#!/bin/bash
# round 2, call H: launch list of the encoder with the fused MLP kernel (per-kernel times + DRAM bytes), then a --set full capture of the stage-1 launch
mkdir -p gpurun_out
python profiles/run_profile.py --iters 2 --max-len 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 220 --csv --log-file gpurun_out/launches_r2h.csv python profiles/run_profile.py --iters 2 --max-len 2 > gpurun_out/prof_ncu.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_r2h.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ki,vi,mi,ii=(h.index(x) for x in ('Kernel Name','Metric Value','Metric Name','ID'))
d={}
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    e=d.setdefault(r[ii],{'name':r[ki][:60]}); e[r[mi]]=r[vi]
for k,e in d.items():
    if 'swin_mlp' in e['name'] or 'layernorm' in e['name']: print(k,e)
PY
ncu --set full --clock-control none --import-source on -k regex:swin_mlp -s 2 -c 1 -o gpurun_out/r2h_swin_mlp python profiles/run_profile.py --iters 2 --max-len 2 > gpurun_out/prof_ncu2.log 2>&1; echo "full capture rc=$?"
