#!/bin/bash
# round 3, call A: the chunked host entry (parity under every plan, restart inside the pipeline) and the vector Q'b sweep; timings per plan;
# cuBLAS at the bulk shapes
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_host_pipeline.py tests/test_gpu_parity.py -m gpu -x -q --timeout 150 --timeout-method=thread > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/a_pytest.log
timeout 200 python tools/r3_e2e.py > gpurun_out/a_e2e.log 2>&1; echo "e2e rc=$?"; tail -40 gpurun_out/a_e2e.log
timeout 200 python tools/r3_solve.py > gpurun_out/a_solve.log 2>&1; echo "solve rc=$?"; tail -30 gpurun_out/a_solve.log
timeout 60 ./build/dmma_rate > gpurun_out/a_dmma.log 2>&1; echo "dmma rc=$?"; cat gpurun_out/a_dmma.log
