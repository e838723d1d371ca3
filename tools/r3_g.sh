#!/bin/bash
# round 3, call G: the pipelined host entry when the CPU wrote the pinned buffer last (bench.py's e2e leg) vs a device-to-host copy
mkdir -p gpurun_out
timeout 150 python tools/r3_plans5.py > gpurun_out/g_plans5.log 2>&1; echo "plans5 rc=$?"; grep -v "step " gpurun_out/g_plans5.log | tail -60
