#!/bin/bash
# call J: wide tests (k_trecon), all tests, state (timeline with hp2 + trecon), bench, config 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/j_all.log 2>&1; echo "all rc=$?"; tail -6 gpurun_out/i_all.log
timeout 300 python tools/r2_state.py 2>&1 | tee gpurun_out/j_state.log | head -12 | cut -c1-1200
timeout 600 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/j_bench.json; tail -3 gpurun_out/j_bench.err
