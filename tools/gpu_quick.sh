#!/bin/bash
# usage: quick.sh tag  -> gpu tests + full-size bench summary
tag=$1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json;d=json.load(open('gpurun_out/bench_$tag.json'));print('RESULT',d['value'],d['ms_per_step'],d['roofline']['groups_ms_per_step'],d['collisions_per_step'])"
tail -3 gpurun_out/bench_$tag.err
