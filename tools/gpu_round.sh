#!/bin/bash
# end-of-milestone GPU run: tests, smoke, bench (ours + reference arm), ncu launch list and full capture
tag=$1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -2 gpurun_out/bench_ref_$tag.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 125 -c 60 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"advect_lean_kernel|scatter_kernel|cell_collide_kernel|advect_full_kernel" -s 8 -c 4 -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -8
