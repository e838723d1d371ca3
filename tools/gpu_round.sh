#!/bin/bash
# One GPU-box round: tests, smoke, bench (both arms), ncu launch list + full captures of the top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 150 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
if [ "$1" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches.csv python tools/prof_one.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_cvy -s 8 -c 1 -f -o gpurun_out/prof_cvy python tools/prof_one.py > gpurun_out/ncu_cvy.log 2>&1; echo "ncu cvy rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_vta -s 8 -c 1 -f -o gpurun_out/prof_vta python tools/prof_one.py > gpurun_out/ncu_vta.log 2>&1; echo "ncu vta rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_panel -s 0 -c 1 -f -o gpurun_out/prof_panel python tools/prof_one.py > gpurun_out/ncu_panel.log 2>&1; echo "ncu panel rc=$?"
fi
