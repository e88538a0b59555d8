#!/bin/bash
mkdir -p gpurun_out
BLAST_CONDUCTOR_DEBUG=1 timeout 300 python tools/bench_render.py --skip-c4 --only-seq 2>&1 | tail -12
CMD="python tools/bench_render.py --skip-c4 --only-seq --voices 1024 --frames 262144"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/seq_launches.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/seq_launches.csv')))
for i,r in enumerate(rows):
    if r and r[0]=="ID": hi=i; break
hdr=rows[hi]; ix={n:j for j,n in enumerate(hdr)}
agg=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>ix['Metric Value']:
        agg[r[ix['Kernel Name']].split('(')[0][-40:]].append(float(r[ix['Metric Value']].replace(',','')))
for k,v in agg.items(): print(f"{k:42s} n={len(v):3d} avg={sum(v)/len(v)/1e3:9.1f} us  total={sum(v)/1e6:8.2f} ms")
PY
