#!/bin/bash
# usage: gpu_final.sh tag -> the round-end single-GPU record: gpu_round.sh (tests, smoke, C5 bench + reference arm, ncu launch
# list + full capture), then the other workloads of BASELINE.json and the 1e9-particle point
tag=$1
bash tools/gpu_round.sh $tag
for wk in c1 c1-grid c2 c4; do
  timeout 900 python bench.py --workload $wk --steps 10 --warmup 3 > gpurun_out/bench_${wk}_$tag.json 2> gpurun_out/bench_${wk}_$tag.err; tail -2 gpurun_out/bench_${wk}_$tag.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${wk}_$tag.json'))
    print('$wk', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['groups_ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'coll', d.get('collisions_per_step'))
except Exception as e:
    print('$wk FAILED', e)
PY
done
timeout 1200 python bench.py --workload c5-1e9 --steps 5 --warmup 3 > gpurun_out/bench_1e9_$tag.json 2> gpurun_out/bench_1e9_$tag.err; tail -2 gpurun_out/bench_1e9_$tag.err
python -c "
import json;d=json.load(open('gpurun_out/bench_1e9_$tag.json'));print('1e9', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['groups_ms_per_step'], d['config'].get('device_memory_gb'))"
