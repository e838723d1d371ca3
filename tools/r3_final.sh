#!/bin/bash
# round 3, final single-GPU validation: every GPU test, smoke, the three bench lines, solve timing, ncu captures of the new kernels
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/z_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/z_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/z_bench.json; tail -2 gpurun_out/z_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/z_bench_ref.json
timeout 300 python bench.py --config 2 > gpurun_out/z_bench_c2.json 2> gpurun_out/z_bench_c2.err; echo "bench c2 rc=$?"; cut -c1-250 gpurun_out/z_bench_c2.json
timeout 200 python tools/r3_solve.py > gpurun_out/z_solve.log 2>&1; echo "solve rc=$?"; tail -14 gpurun_out/z_solve.log
timeout 200 python tools/r3_e2e.py quick > gpurun_out/z_e2e.log 2>&1; echo "e2e rc=$?"; grep -v "step " gpurun_out/z_e2e.log | tail -14
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_qt_dot -s 2 -c 1 -f -o gpurun_out/prof_qtdot python tools/prof_solve.py > gpurun_out/ncu_qtdot.log 2>&1; echo "ncu qt_dot rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_qt_axpy -s 2 -c 1 -f -o gpurun_out/prof_qtaxpy python tools/prof_solve.py > gpurun_out/ncu_qtaxpy.log 2>&1; echo "ncu qt_axpy rc=$?"
ls -la gpurun_out/*.ncu-rep
timeout 200 python tools/r3_plans2.py > gpurun_out/z_plans2.log 2>&1; echo "plans2 rc=$?"; cat gpurun_out/z_plans2.log
