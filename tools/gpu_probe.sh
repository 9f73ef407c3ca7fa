#!/bin/bash
# one-off GPU diagnostics: MN-major probe, seeds debug, phase timing of the tile kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 nice_slam_b200/csrc/probe/probe_mn32c > gpurun_out/probe_mn32c.log 2>&1; echo "probe exit $?" >> gpurun_out/probe_mn32c.log
timeout 300 python tools/debug_seeds.py > gpurun_out/debug_seeds.log 2>&1; echo "exit $?" >> gpurun_out/debug_seeds.log
NSB_LIB=nice_slam_b200/libnsb_timing.so timeout 300 python tools/phase_timing.py 200 > gpurun_out/phase_tile_200.txt 2>&1
NSB_LIB=nice_slam_b200/libnsb_timing.so timeout 300 python tools/phase_timing.py 8192 > gpurun_out/phase_tile_8192.txt 2>&1
cat gpurun_out/probe_mn32c.log gpurun_out/debug_seeds.log gpurun_out/phase_tile_200.txt gpurun_out/phase_tile_8192.txt
