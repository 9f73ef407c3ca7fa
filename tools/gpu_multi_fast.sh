#!/bin/bash
# quick multi-GPU check: the 2-GPU parity tests + the headline part of the bench at N ranks (NSB_BENCH_FAST=1: no mapping / scene extras)
tag=${1:-mf}; n=${2:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peers_single.py -q -m gpu 2>&1 | tail -30 > gpurun_out/${tag}_tests.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests.log; tail -4 gpurun_out/${tag}_tests.log
NSB_BENCH_FAST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 300 --warmup 10 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
echo "bench n=$n exit $?"
python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n ms/step", round(d["ms_per_step"],5), "rays/s", round(d["value"]), "e2e", d["e2e"], "launches", d["gpu_launches"])
except Exception as e:
    print("parse failed", e)
PYEOF
tail -5 gpurun_out/${tag}_bench_n$n.err
