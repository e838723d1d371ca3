#!/bin/bash
# round 2: link bandwidth + host-entry stage trace, then the ncu launch list and full captures of the hot kernels
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 200 python tools/r2_pcie.py > gpurun_out/g_pcie.log 2>&1; echo "pcie rc=$?"; cat gpurun_out/g_pcie.log | tail -30
timeout 300 python -m pytest tests/test_gpu_wide.py -m gpu -q --timeout 200 --timeout-method=thread 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python tools/prof_one.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_cvy_p -s 1 -c 1 -f -o gpurun_out/prof_cvy python tools/prof_one.py > gpurun_out/ncu_cvy.log 2>&1; echo "ncu cvy rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_vta -s 2 -c 1 -f -o gpurun_out/prof_vta python tools/prof_one.py > gpurun_out/ncu_vta.log 2>&1; echo "ncu vta rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vpk_rmul -s 0 -c 1 -f -o gpurun_out/prof_rmul python tools/prof_one.py > gpurun_out/ncu_rmul.log 2>&1; echo "ncu rmul rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chol128 -s 0 -c 1 -f -o gpurun_out/prof_chol python tools/prof_one.py > gpurun_out/ncu_chol.log 2>&1; echo "ncu chol rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_hr128 -s 0 -c 1 -f -o gpurun_out/prof_hr python tools/prof_one.py > gpurun_out/ncu_hr.log 2>&1; echo "ncu hr rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_apply1_tma -s 1 -c 1 -f -o gpurun_out/prof_apply1 python tools/prof_one.py 8192 1024 1 > gpurun_out/ncu_apply1.log 2>&1; echo "ncu apply1 rc=$?"
ls -la gpurun_out/*.ncu-rep
