#!/bin/bash
# round 3, call B: stage timeline of the pipelined host entry; the vector Q'b kernels after their rewrite (tests + timing)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_host_pipeline.py -m gpu -x -q --timeout 150 --timeout-method=thread -k "vector or plan" > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/b_pytest.log
timeout 200 python tools/r3_e2e.py quick > gpurun_out/b_e2e.log 2>&1; echo "e2e rc=$?"; tail -70 gpurun_out/b_e2e.log
timeout 200 python tools/r3_solve.py > gpurun_out/b_solve.log 2>&1; echo "solve rc=$?"; tail -30 gpurun_out/b_solve.log
