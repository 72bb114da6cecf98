#!/bin/bash
# usage: gpu_n.sh tag N [extra bench args] -> multi-rank tests + N-GPU bench line (no e2e)
tag=$1; n=$2; shift 2
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_exchange_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/bench_n${n}_$tag.json 2> gpurun_out/bench_n${n}_$tag.err
tail -3 gpurun_out/bench_n${n}_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n${n}_$tag.json'))
print('RESULT', d['value'], d['ms_per_step'], d['roofline']['groups_ms_per_step'])
for r in d['per_rank']: print(r)
PY
