#!/bin/bash
# like gpu_sweep.sh, at the full bench size (1.25e8 particles, 3.1e6 tets): ~80 s per variant (setup dominates)
for so in "$@"; do
  PANTERA_GPU_LIB=$PWD/$so timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > /tmp/sw.json 2> /tmp/sw.err
  python -c "
import json;d=json.load(open('/tmp/sw.json'));g=d['roofline']['groups_ms_per_step'];print('SWEEP $so', round(d['ms_per_step'],3), {k:round(v,3) for k,v in g.items()})" || tail -3 /tmp/sw.err
done
