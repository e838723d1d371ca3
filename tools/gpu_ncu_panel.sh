#!/bin/bash
# ncu launch list of one qr! + one full capture of the panel kernel (source-level), for profiles/.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches.csv python tools/prof_one.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_panel -s 0 -c 1 -f -o gpurun_out/prof_panel python tools/prof_one.py > gpurun_out/ncu_panel.log 2>&1; echo "ncu panel rc=$?"
