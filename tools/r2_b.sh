#!/bin/bash
# round 2, call C: wide chain v2 (register-resident Cholesky / LU) parity + timing, new tests, new bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py -m gpu -q -x --timeout 200 --timeout-method=thread > gpurun_out/c_wide.log 2>&1; echo "wide rc=$?"; tail -15 gpurun_out/b_wide.log
timeout 300 python tools/r2_state.py 2>&1 | tee gpurun_out/c_state.log | tail -30
timeout 900 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread --deselect tests/test_gpu_wide.py > gpurun_out/c_all.log 2>&1; echo "all rc=$?"; tail -8 gpurun_out/b_all.log
timeout 600 python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"; cat gpurun_out/c_bench.json; tail -3 gpurun_out/c_bench.err
timeout 300 python bench.py --config 2 --no-cpu > gpurun_out/c_bench_c2.json 2> gpurun_out/c_bench_c2.err; echo "bench c2 rc=$?"; cat gpurun_out/c_bench_c2.json; tail -3 gpurun_out/c_bench_c2.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/c_bench_ref.json
