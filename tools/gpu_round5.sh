#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/${tag}_tests.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests.log; tail -5 gpurun_out/${tag}_tests.log
NSB_SMALL_RAYS=0 timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/${tag}_tests_small0.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests_small0.log; tail -5 gpurun_out/${tag}_tests_small0.log
NSB_MLP_BACKEND=2 timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/${tag}_tests_group.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests_group.log; tail -5 gpurun_out/${tag}_tests_group.log
for v in "NSB_X=0" "NSB_SMALL_RAYS=0"; do
  env $v NSB_BENCH_FAST=1 timeout 600 python bench.py --steps 300 --warmup 10 > gpurun_out/${tag}_fast_${v%%=*}.json 2>/dev/null
  python - <<PYEOF
import json
d=json.load(open("gpurun_out/${tag}_fast_${v%%=*}.json"))
print("fast bench [$v]: ms/step", round(d["ms_per_step"],5), "e2e", round(d["e2e"]["ms_per_step"],5), "warm", round(d["extra"]["l2_warm_ms_per_step"],5), "bwd", round(d["roofline"]["launch_ms"],5))
PYEOF
done
NSB_SMALL_RAYS=0 NSB_LIB=nice_slam_b200/libnsb_timing.so timeout 300 python tools/phase_timing.py 200 > gpurun_out/${tag}_phase_200.txt 2>&1; cat gpurun_out/${tag}_phase_200.txt
timeout 1200 python bench.py --steps 300 --warmup 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; python - <<PYEOF
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "rays/s", d["value"], "e2e", d["e2e"]["ms_per_step"], "warm", d["extra"]["l2_warm_ms_per_step"])
for k in ("mapping_configs1","mapping_loop_step"): print(k, d["extra"][k]["ms_per_step"])
print("dropin", d["extra"].get("dropin"))
print([ (x["rays"], x["samples"], round(x["ms_per_step"],3), round(x["rays_per_s"]/1e6,2)) for x in d["extra"]["sweep_tracking_iteration"]])
print([(x["scene"], round(x["ms_per_step"],3)) for x in d["extra"]["mapping_other_scenes"]])
PYEOF
tail -3 gpurun_out/${tag}_bench.err
