#!/bin/bash
# round 3, call D: width of the first upload of the pipelined host entry; pipeline tests on the new default plan
mkdir -p gpurun_out
timeout 200 python tools/r3_plans3.py > gpurun_out/d_plans3.log 2>&1; echo "plans3 rc=$?"; grep -v "step " gpurun_out/d_plans3.log | tail -40
timeout 300 python -m pytest tests/test_gpu_host_pipeline.py tests/test_gpu_parity.py -m gpu -x -q --timeout 150 --timeout-method=thread -k "plan or pipeline or narrower or host" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d_pytest.log
