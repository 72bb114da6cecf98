#!/bin/bash
tag=$1
timeout 900 python -m pytest tests/test_gpu_advect.py tests/test_gpu_exchange_parity.py tests/test_gpu_collide.py tests/test_gpu_scale_parity.py -m gpu -x -q 2>&1 | tail -3
for wk in c1 c4; do
  timeout 600 python bench.py --workload $wk --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${wk}_$tag.json 2> gpurun_out/bench_${wk}_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_${wk}_$tag.json'));print('$wk', d['value'], d['ms_per_step'], d['roofline']['groups_ms_per_step'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cell_collide_kernel" -s 3 -c 1 -o gpurun_out/prof_c2_$tag python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu --particles 20000000 > gpurun_out/ncu_c2_$tag.log 2>&1
tail -2 gpurun_out/ncu_c2_$tag.log | cut -c1-300
