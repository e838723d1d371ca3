#!/bin/bash
# 2-GPU call: NCCL parity tests (incl. the restart after a refused panel), bench at N=2 with the even and the balanced split
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 200 --timeout-method=thread > gpurun_out/n2_tests.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/n2_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/n2_bench.json; tail -2 gpurun_out/n2_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --split balanced --no-e2e --no-solve > gpurun_out/n2_bench_bal.json 2> gpurun_out/n2_bench_bal.err; echo "bench balanced rc=$?"; cut -c1-300 gpurun_out/n2_bench_bal.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/n2_bench_ref.json 2> gpurun_out/n2_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/n2_bench_ref.json
