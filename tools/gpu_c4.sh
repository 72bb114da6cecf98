#!/bin/bash
tag=$1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E    |passed|failed" | cut -c1-300 | head -12
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c4_$tag.json 2> gpurun_out/bench_c4_$tag.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c4_$tag.json'));print('c4', d['value'], d['ms_per_step'], d['roofline']['groups_ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 70 --csv --log-file gpurun_out/launches_c4_$tag.csv python bench.py --workload c4 --steps 3 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python - <<P
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_c4_$tag.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4][:70]; v=float(r[-1].replace(',','')); u=r[-2]
    if u=='us': v/=1e3
    elif u=='ns': v/=1e6
    agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in sorted(agg.items(),key=lambda kv:-kv[1][1])[:14]: print(f"{t:8.3f} ms {n:4d}  {k}")
P
