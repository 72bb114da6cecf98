#!/bin/bash
# usage: gpu_ab.sh tag so1 so2 ...  -> full-size sweep of the builds, then ncu full capture of advect_lean of the default build
tag=$1; shift
bash tools/gpu_sweep_full.sh "$@"
bash tools/gpu_ncu_one.sh $tag "advect_lean_kernel"
