#!/bin/bash
# programmatic dependent launch of the backward: GPU suite + headline iteration with the option off / on
tag=${1:-p}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
[ -n "$NOTESTS" ] || timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/${tag}_tests.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests.log; tail -5 gpurun_out/${tag}_tests.log
for v in "NSB_PDL=0" "NSB_PDL=1"; do
  env $v NSB_BENCH_FAST=1 timeout 600 python bench.py --steps 300 --warmup 10 > gpurun_out/${tag}_fast_${v#*=}.json 2>gpurun_out/${tag}_fast_${v#*=}.err
  python - <<PYEOF
import json
d=json.load(open("gpurun_out/${tag}_fast_${v#*=}.json"))
print("fast bench [$v]: ms/step", round(d["ms_per_step"],5), "e2e", round(d["e2e"]["ms_per_step"],5), "warm", round(d["extra"]["l2_warm_ms_per_step"],5), "bwd", round(d["roofline"]["launch_ms"],5), "f16", d["extra"].get("fwd_f16_option",{}).get("ms_per_step"))
PYEOF
done
