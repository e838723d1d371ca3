#!/bin/bash
# full captures of the BULK launches of step 0 (launch order: chain apply, second apply, bulk)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_cvy_p -s 2 -c 1 -f -o gpurun_out/prof_cvy python tools/prof_one.py > gpurun_out/ncu_cvy.log 2>&1; echo "ncu cvy rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_vta -s 6 -c 1 -f -o gpurun_out/prof_vta python tools/prof_one.py > gpurun_out/ncu_vta.log 2>&1; echo "ncu vta rc=$?"
