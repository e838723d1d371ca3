#!/bin/bash
# round 2, call A: wide-panel chain parity first, then the whole GPU suite, then timing (default / wide off) and a serial profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py -m gpu -q --timeout 200 --timeout-method=thread > gpurun_out/a_wide.log 2>&1; echo "wide rc=$?"; tail -30 gpurun_out/a_wide.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread --deselect tests/test_gpu_wide.py > gpurun_out/a_all.log 2>&1; echo "all rc=$?"; tail -8 gpurun_out/a_all.log
timeout 300 python tools/r2_state.py 2>&1 | tee gpurun_out/a_state.log | tail -40
