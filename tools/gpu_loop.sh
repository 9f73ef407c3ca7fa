#!/bin/bash
tag=${1:-l}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ba.py -q -m gpu -k "loop or mapper or coarse or backends or weight_gradients or fp16 or golden or render" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log; echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/${tag}_tests.log; tail -5 gpurun_out/${tag}_tests.log
timeout 300 python tools/loop_step.py 2>&1 | tail -4
