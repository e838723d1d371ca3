"""Quick tuning sweep on the bench workload: panel CTA count / nb, with the per-class profile."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dhqr_b200 as D
dev = torch.device("cuda:0"); h = D.default_handle(0)
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 4096)
A = D.colmajor_empty(m, n, dev); al = torch.zeros(n, dtype=torch.float64, device=dev)
fl = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3
def timeit(nb=0, reps=3):
    best = 1e30
    for _ in range(reps + 1):
        D.fill_uniform_(A, 0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.householder_(A, al, nb); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def profile(nb=0):
    h.set_option("profile", 1); D.fill_uniform_(A, 0); torch.cuda.synchronize(); h.profile_reset()
    D.householder_(A, al, nb); torch.cuda.synchronize(); p = h.profile(); h.set_option("profile", 0)
    return {k: (round(v["ms"], 2), v["count"], round(v["work"] / v["ms"] / 1e9, 1) if k.startswith("k_gemm") and v["ms"] > 0 else None) for k, v in p.items()}
for pc in [0, 148, 128, 96, 74, 64, 48, 32]:
    h.set_option("panel_ctas", pc)
    t = timeit()
    print(f"panel_ctas={pc}: {t:.2f} ms  {fl / t / 1e9:.2f} TFLOP/s", flush=True)
h.set_option("panel_ctas", 0)
print("profile nb=128:", json.dumps(profile()), flush=True)
for nb in [32, 64, 96]:
    t = timeit(nb)
    print(f"nb={nb}: {t:.2f} ms  {fl / t / 1e9:.2f} TFLOP/s", flush=True)
print("profile nb=64:", json.dumps(profile(64)), flush=True)
