#!/bin/bash
# probe + parity tests with both weight-gradient paths + phase timing + short bench
tag=${1:-q}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for c in $(seq 0 15); do timeout 60 nice_slam_b200/csrc/probe/probe_mn32c $c 2>&1 | grep -E "combo|status" ; done > gpurun_out/${tag}_probe_mn32c.log 2>&1
cat gpurun_out/${tag}_probe_mn32c.log
NSB_WGRAD_TC=0 timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/${tag}_tests_fp32wg.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests_fp32wg.log
tail -12 gpurun_out/${tag}_tests_fp32wg.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -120 > gpurun_out/${tag}_tests.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests.log
tail -40 gpurun_out/${tag}_tests.log
NSB_LIB=nice_slam_b200/libnsb_timing.so timeout 300 python tools/phase_timing.py 200 > gpurun_out/${tag}_phase_200.txt 2>&1
NSB_LIB=nice_slam_b200/libnsb_timing.so timeout 300 python tools/phase_timing.py 8192 > gpurun_out/${tag}_phase_8192.txt 2>&1
cat gpurun_out/${tag}_phase_200.txt
timeout 1200 python bench.py --steps 200 --warmup 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; python - <<PYEOF
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "rays/s", d["value"], "e2e", d["e2e"]["ms_per_step"], "warm", d["extra"]["l2_warm_ms_per_step"])
print("bwd launch ms", d["roofline"]["launch_ms"], "iteration_frac", d["roofline"].get("iteration_frac"), "tensor frac", d.get("roofline_tensor",{}).get("frac"))
for k in ("mapping_configs1","mapping_loop_step"): print(k, d["extra"][k]["ms_per_step"])
print("dropin", d["extra"].get("dropin"))
print([ (x["rays"], x["samples"], round(x["ms_per_step"],3), round(x["rays_per_s"]/1e6,2)) for x in d["extra"]["sweep_tracking_iteration"]])
print([(x["scene"], round(x["ms_per_step"],3)) for x in d["extra"]["mapping_other_scenes"]])
print(d["extra"]["mapping_sharded_masked"])
PYEOF
tail -5 gpurun_out/${tag}_bench.err
