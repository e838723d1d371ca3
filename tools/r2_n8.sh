#!/bin/bash
# 8-GPU call: scaling series N = 8, 4 on the bench workload, BASELINE config 5 (65536 x 8192: qr! + H \ b + residuals)
mkdir -p gpurun_out
for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench N=$N rc=$?"; cut -c1-400 gpurun_out/n${N}_bench.json; tail -2 gpurun_out/n${N}_bench.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 tools/dist_config.py 65536 8192 > gpurun_out/n8_config5.log 2>&1; echo "config5 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/n8_config5.log | tail -6
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 tools/r2_timeline.py > gpurun_out/n8_timeline.log 2>&1; echo "timeline rc=$?"; tail -8 gpurun_out/n8_timeline.log
