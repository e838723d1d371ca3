#!/bin/bash
# round 2, call F: every GPU test (wide chain v3, complex path, Q apply, persistent cvy, new tinv), state, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/f_all.log 2>&1; echo "all rc=$?"; tail -12 gpurun_out/d_all.log
timeout 300 python tools/r2_state.py 2>&1 | tee gpurun_out/f_state.log | tail -30
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; cat gpurun_out/f_bench.json; tail -3 gpurun_out/f_bench.err
timeout 300 python bench.py --config 2 --no-cpu > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err; echo "bench c2 rc=$?"; cat gpurun_out/f_bench_c2.json; tail -3 gpurun_out/f_bench_c2.err
