#!/bin/bash
# multi-GPU evidence: 2-GPU parity test + bench at N ranks (torchrun), N = $2 (default 2)
tag=${1:-m}; n=${2:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_smi.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peers_single.py -q -m gpu 2>&1 | tail -30 > gpurun_out/${tag}_tests.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests.log; tail -6 gpurun_out/${tag}_tests.log
for k in ${SKIP1:+} $([ -n "$SKIP1" ] || echo 1) $n; do
  if [ $k -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 300 --warmup 10 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $k --steps 300 --warmup 10 > gpurun_out/${tag}_bench_n$k.json 2> gpurun_out/${tag}_bench_n$k.err
  fi
  echo "bench n=$k exit $?"
  python - <<PYEOF
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_n$k.json").read().strip().splitlines()[-1])
    print("N=$k ms/step", round(d["ms_per_step"],5), "rays/s", round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],5), "launches", d["gpu_launches"], d["run"]["exchange"])
    m=d["extra"]["mapping_sharded_masked"]; print("  sharded mapping:", round(m["ms_per_step"],4), "collectives", m["collectives_per_step"], m["launch"])
    print("  scenes:", [(x["scene"], round(x["ms_per_step"],3), x["collectives_per_step"]) for x in d["extra"]["mapping_other_scenes"]])
except Exception as e:
    print("parse failed", e)
PYEOF
  tail -3 gpurun_out/${tag}_bench_n$k.err
done
