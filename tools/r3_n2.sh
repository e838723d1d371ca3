#!/bin/bash
# round 3, 2-GPU call: NCCL parity tests (per-panel T' slots, Q'b with T' prepared before b arrives), bench at N=2 (solve block)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 200 --timeout-method=thread > gpurun_out/n2_tests.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/n2_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open("gpurun_out/n2_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["ms_per_step"], "solve", {k:round(v,3) for k,v in d["solve"].items() if k.endswith("_ms")}, "parity", d["parity"])
P
tail -2 gpurun_out/n2_bench.err
