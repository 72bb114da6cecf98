#!/bin/bash
tag=$1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
nvidia-smi --query-gpu=memory.total,memory.used --format=csv; free -g | head -2
timeout 1700 python bench.py --workload c5-1e9 --steps 5 --warmup 3 > gpurun_out/bench_c5-1e9_$tag.json 2> gpurun_out/bench_c5-1e9_$tag.err; tail -3 gpurun_out/bench_c5-1e9_$tag.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_c5-1e9_$tag.json'))
    print('1e9', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['groups_ms_per_step'], d['config'])
except Exception as e:
    print('1e9 FAILED', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"advect_lean_kernel" -s 3 -c 1 -o gpurun_out/prof_c4_$tag python bench.py --workload c4 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4_$tag.log 2>&1
ls -la gpurun_out/prof_c4_$tag.ncu-rep
