#!/bin/bash
tag=$1
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_c4_$tag.csv python bench.py --workload c4 --steps 4 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python - <<PY
import csv, io
from collections import OrderedDict
lines=[l for l in open('gpurun_out/launches_c4_$tag.csv') if l.startswith('"')]
rows=list(csv.DictReader(io.StringIO("".join(lines))))
agg=OrderedDict()
for r in rows:
    k=r["Kernel Name"].split("(")[0][:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r["Metric Value"])/1e6
for k,(n,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{n:4d} {ms:8.3f} ms  {ms/n:7.3f} ms/launch  {k}")
PY
