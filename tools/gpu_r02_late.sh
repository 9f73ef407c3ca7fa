#!/bin/bash
# late-round-2 validation on one GPU: smoke(), the GPU suite, the full bench line [, the launch list of one mapping iteration]
#   tools/gpu_r02_late.sh <tag> [ncu]
tag=${1:-r02late}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/${tag}_tests.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests.log; tail -6 gpurun_out/${tag}_tests.log
timeout 1200 python bench.py --steps 300 --warmup 10 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; python - <<PYEOF
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "rays/s", d["value"], "e2e", d["e2e"], "warm", d["extra"]["l2_warm_ms_per_step"])
for k in ("mapping_configs1","mapping_loop_step","mapping_loop_step_coarse_mapper"): print(k, d["extra"][k]["ms_per_step"])
print([(x["scene"], round(x["ms_per_step"],3)) for x in d["extra"]["mapping_other_scenes"]])
print("cpu", d["cpu_baseline"])
PYEOF
tail -3 gpurun_out/${tag}_bench.err
if [ "$2" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_map_launches.csv python tools/map_launches.py 996 > gpurun_out/${tag}_map_launches.log 2>&1
python - <<PYEOF
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_map_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki].split("(")[0], float(r[vi])) for r in rows[1:]]
idx=[i for i,(k,_) in enumerate(seq) if "render_fwd" in k]
for k,v in seq[idx[-1]:]: print("%-60s %10.1f us" % (k[:60], v/1000 if v>1000 else v))
PYEOF
fi
