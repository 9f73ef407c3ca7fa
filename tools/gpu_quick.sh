#!/bin/bash
# quick GPU check: parity tests (stop at first failure), phase timing, short bench
tag=${1:-q}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 > gpurun_out/${tag}_tests.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests.log
tail -6 gpurun_out/${tag}_tests.log
NSB_LIB=nice_slam_b200/libnsb_timing.so timeout 300 python tools/phase_timing.py 200 > gpurun_out/${tag}_phase_200.txt 2>&1
NSB_LIB=nice_slam_b200/libnsb_timing.so timeout 300 python tools/phase_timing.py 8192 > gpurun_out/${tag}_phase_8192.txt 2>&1
cat gpurun_out/${tag}_phase_200.txt
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; python - <<PYEOF
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "rays/s", d["value"], "e2e", d["e2e"]["ms_per_step"], "warm", d["extra"]["l2_warm_ms_per_step"])
print("bwd launch ms", d["roofline"]["launch_ms"])
for k in ("mapping_configs1","mapping_loop_step"): print(k, d["extra"][k]["ms_per_step"])
print([ (x["rays"], round(x["ms_per_step"],3), round(x["rays_per_s"]/1e6,2)) for x in d["extra"]["sweep_tracking_iteration"]])
PYEOF
tail -3 gpurun_out/${tag}_bench.err
