#!/bin/bash
# final single-GPU validation of round 2: every GPU test, smoke, both bench arms, config 2, state, ncu captures of the bulk GEMMs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/z_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/z_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/z_bench.json; tail -2 gpurun_out/z_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/z_bench_ref.json
timeout 300 python bench.py --config 2 > gpurun_out/z_bench_c2.json 2> gpurun_out/z_bench_c2.err; echo "bench c2 rc=$?"; cut -c1-250 gpurun_out/z_bench_c2.json
timeout 300 python tools/r2_state.py > gpurun_out/z_state.log 2>&1; head -8 gpurun_out/z_state.log | cut -c1-700
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_cvy_p -s 2 -c 1 -f -o gpurun_out/prof_cvy python tools/prof_one.py > gpurun_out/ncu_cvy.log 2>&1; echo "ncu cvy rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_vta -s 6 -c 1 -f -o gpurun_out/prof_vta python tools/prof_one.py > gpurun_out/ncu_vta.log 2>&1; echo "ncu vta rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python tools/prof_one.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:k_unblocked_wave -c 1 -f -o gpurun_out/prof_wave python tools/prof_one.py 8192 1024 1 > gpurun_out/ncu_wave.log 2>&1; echo "ncu wave rc=$?"
