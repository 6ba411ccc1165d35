#!/bin/bash
mkdir -p gpurun_out
python bench.py --config smoke128x64-ddim --steps 2 --warmup 2 --profile > gpurun_out/r2_prof_ddim128.json 2> gpurun_out/r2_prof_ddim128.err
grep " ms " gpurun_out/r2_prof_ddim128.err | head -30
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_jellyfish.csv python bench.py --config jellyfish128 --steps 1 --warmup 1 > gpurun_out/r2_ncu_jelly.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_jellyfish.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e6
    elif r[ui]=='us': v/=1e3
    agg[r[ki][:90]][0]+=1; agg[r[ki][:90]][1]+=v
tot=sum(v for _,v in agg.values())
for k,(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]: print(f"{v:9.2f} ms {100*v/tot:5.1f}% n={n:4d} {k}")
print('total',tot)
PY
