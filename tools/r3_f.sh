#!/bin/bash
# round 3, call F: does the Gram kernel of the panel chain interact with the pipelined host entry?  (bench e2e 54.6 ms with k_gram_sym)
mkdir -p gpurun_out
timeout 200 python tools/r3_plans4.py > gpurun_out/f_plans4.log 2>&1; echo "plans4 rc=$?"; grep -v "step " gpurun_out/f_plans4.log | tail -45
