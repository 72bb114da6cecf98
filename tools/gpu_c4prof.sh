#!/bin/bash
tag=$1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for wk in c1 c4; do
  timeout 600 python bench.py --workload $wk --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${wk}_$tag.json 2> gpurun_out/bench_${wk}_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_${wk}_$tag.json'));print('$wk', d['value'], d['ms_per_step'], d['roofline']['groups_ms_per_step'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mcc_vahedi_kernel|mcc_select" -s 8 -c 2 -o gpurun_out/prof_c4m_$tag python bench.py --workload c4 --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c4m_$tag.log 2>&1
tail -2 gpurun_out/ncu_c4m_$tag.log | cut -c1-300
