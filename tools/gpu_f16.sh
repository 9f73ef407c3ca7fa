#!/bin/bash
# option fwd_f16 (FP16 hi|lo forward): its own parity test, the whole GPU suite with the option on, and the headline iteration both ways
tag=${1:-f}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_peers_single.py -q -m gpu -x -k "fp16_split or peers" 2>&1 | tail -30 > gpurun_out/${tag}_tests_f16.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests_f16.log; tail -12 gpurun_out/${tag}_tests_f16.log
NSB_FWD_F16=1 timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/${tag}_tests_all_f16.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests_all_f16.log; tail -8 gpurun_out/${tag}_tests_all_f16.log
for v in "NSB_FWD_F16=0" "NSB_FWD_F16=1"; do
  env $v NSB_BENCH_FAST=1 timeout 600 python bench.py --steps 300 --warmup 10 > gpurun_out/${tag}_fast_${v#*=}.json 2>gpurun_out/${tag}_fast_${v#*=}.err
  python - <<PYEOF
import json
d=json.load(open("gpurun_out/${tag}_fast_${v#*=}.json"))
print("fast bench [$v]: ms/step", round(d["ms_per_step"],5), "e2e", round(d["e2e"]["ms_per_step"],5), "warm", round(d["extra"]["l2_warm_ms_per_step"],5), "bwd", round(d["roofline"]["launch_ms"],5))
PYEOF
done
