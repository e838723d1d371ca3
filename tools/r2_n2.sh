#!/bin/bash
# 2-GPU call: NCCL parity tests (incl. the restart after a refused panel), bench at N=2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 200 --timeout-method=thread > gpurun_out/n2_tests.log 2>&1; echo "dist tests rc=$?"; tail -5 gpurun_out/n2_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc=$?"; cat gpurun_out/n2_bench.json | cut -c1-3000; tail -3 gpurun_out/n2_bench.err
