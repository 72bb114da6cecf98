#!/bin/bash
# A/B sweep of differently tuned builds of libpantera_gpu.so (developer tool): prints per-group ms per step
for so in "$@"; do
  PANTERA_GPU_LIB=$PWD/$so timeout 300 python bench.py --steps 6 --warmup 3 --particles 31250000 --cubes 8 128 128 --no-cpu --no-e2e > /tmp/sw.json 2> /tmp/sw.err
  python -c "
import json;d=json.load(open('/tmp/sw.json'));g=d['roofline']['groups_ms_per_step'];print('SWEEP $so', round(d['ms_per_step'],3), {k:round(v,3) for k,v in g.items()})" || tail -3 /tmp/sw.err
done
