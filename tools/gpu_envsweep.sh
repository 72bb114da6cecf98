#!/bin/bash
# usage: gpu_envsweep.sh VAR v1 v2 ...  -> full-size bench per value of the environment variable
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > /tmp/sw.json 2> /tmp/sw.err
  python -c "
import json;d=json.load(open('/tmp/sw.json'));g=d['roofline']['groups_ms_per_step'];print('SWEEP $var=$v', round(d['ms_per_step'],3), {k:round(x,3) for k,x in g.items()})" || tail -3 /tmp/sw.err
done
