#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
for c in $(seq 0 15); do timeout 60 nice_slam_b200/csrc/probe/probe_mn32c $c 2>&1 | grep -E "combo|status" ; done > gpurun_out/${tag}_probe_mn32c.log 2>&1
cat gpurun_out/${tag}_probe_mn32c.log
bash tools/gpu_quick.sh $tag
