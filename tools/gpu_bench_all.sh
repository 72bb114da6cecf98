#!/bin/bash
# usage: gpu_bench_all.sh tag -> injection tests, default bench line (C5, e2e, cpu), the reference arm, and the other workloads
tag=$1
timeout 600 python -m pytest tests/test_injection.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('C5', d['value'], d['ms_per_step'], d['roofline']['groups_ms_per_step'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['sequential'])
print('cpu', d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -1 gpurun_out/bench_ref_$tag.err; cut -c1-300 gpurun_out/bench_ref_$tag.json
for wk in c1 c1-grid c2 c4; do
  timeout 900 python bench.py --workload $wk --steps 10 --warmup 3 > gpurun_out/bench_${wk}_$tag.json 2> gpurun_out/bench_${wk}_$tag.err; tail -2 gpurun_out/bench_${wk}_$tag.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${wk}_$tag.json'))
    print('$wk', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['groups_ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'coll', d['collisions_per_step'])
except Exception as e:
    print('$wk FAILED', e)
PY
done
