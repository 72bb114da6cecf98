#!/bin/bash
# usage: gpu_ncu_one.sh tag kernel_regex [extra bench args]  -> one ncu --set full capture of the matching kernels (2 launches after warm-up)
tag=$1; k=$2; shift 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 2 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/ncu_full_$tag.log
ls -la gpurun_out/prof_$tag.ncu-rep
