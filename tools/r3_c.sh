#!/bin/bash
# round 3, call C: planner knobs of the pipelined host entry (step-time assumption, catch-up streams, chunk width) + timeline of the best
mkdir -p gpurun_out
timeout 300 python tools/r3_plans.py > gpurun_out/c_plans.log 2>&1; echo "plans rc=$?"; tail -100 gpurun_out/c_plans.log
timeout 200 python -m pytest tests/test_gpu_host_pipeline.py -m gpu -x -q --timeout 150 --timeout-method=thread -k "plan or pipeline or narrower" > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c_pytest.log
