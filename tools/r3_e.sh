#!/bin/bash
# round 3, call E: k_gram_sym (symmetric Gram of a packed panel) in the wide chain and in the Q'b preparation: parity tests, A/B bench, solve timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/e_pytest.log 2>&1; rc=$?; echo "pytest (gram_sym=1) rc=$rc"; tail -4 gpurun_out/e_pytest.log
if [ $rc -ne 0 ]; then
  DHQR_GRAM_SYM=0 timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_parity.py -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/e_pytest0.log 2>&1; echo "pytest (gram_sym=0) rc=$?"; tail -4 gpurun_out/e_pytest0.log
fi
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"; python - <<'P'
import json
d=json.loads(open("gpurun_out/e_bench.json").read().strip().splitlines()[-1])
print("gram_sym=1:", round(d["ms_per_step"],3), "ms", round(d["value"]), "GFLOP/s; e2e", round(d["e2e"]["ms_per_step"],2), "ms; solve", {k:round(v,3) for k,v in d["solve"].items() if k.endswith("_ms")}, "parity", d["parity"]["qr_residual_fro_rel"])
cl=d["roofline"]["classes"]; print({k:(v["ms"],v["count"]) for k,v in cl.items() if k in ("k_gram128","k_wreduce","k_gram2_finish","k_vpk_rmul","k_pack")})
P
DHQR_GRAM_SYM=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-solve > gpurun_out/e_bench0.json 2> gpurun_out/e_bench0.err; echo "bench0 rc=$?"; python - <<'P'
import json
d=json.loads(open("gpurun_out/e_bench0.json").read().strip().splitlines()[-1])
print("gram_sym=0:", round(d["ms_per_step"],3), "ms", round(d["value"]), "GFLOP/s")
cl=d["roofline"]["classes"]; print({k:(v["ms"],v["count"]) for k,v in cl.items() if k in ("k_gram128","k_wreduce","k_gram2_finish","k_vpk_rmul","k_pack")})
P
timeout 200 python tools/r3_solve.py > gpurun_out/e_solve.log 2>&1; echo "solve rc=$?"; tail -12 gpurun_out/e_solve.log
