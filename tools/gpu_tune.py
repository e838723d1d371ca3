"""Quick tuning sweep on the bench workload: look-ahead on/off, panel CTA count, vta chunk cap."""
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
def resid():
    A0 = D.colmajor_empty(m, n, dev); D.fill_uniform_(A0, 0)
    R = torch.zeros(m, n, dtype=torch.float64, device=dev)
    R[:n] = torch.triu(A[:n], 1) + torch.diag(al)
    for k in range(((n - 1) // 128) * 128, -1, -128):
        kb = min(128, n - k); V = torch.tril(A[k:, k:k + kb])
        Tinv = torch.eye(kb, dtype=torch.float64, device=dev) + torch.triu(V.T @ V, 1)
        R[k:] -= V @ torch.linalg.solve_triangular(Tinv, V.T @ R[k:], upper=True)
    return float(torch.linalg.norm(R - A0) / torch.linalg.norm(A0))
def show(tag):
    t = timeit()
    print(f"{tag}: {t:.2f} ms  {fl / t / 1e9:.2f} TFLOP/s", flush=True)
h.set_option("lookahead", 0); show("serial")
h.set_option("lookahead", 1); show("lookahead default")
print("   resid (lookahead):", resid(), flush=True)
for pc in (96, 64):
    h.set_option("panel_ctas", pc); show(f"lookahead panel_ctas={pc}")
h.set_option("panel_ctas", 0)
for mc in (8, 16, 48, 1000):
    h.set_option("vta_max_chunks", mc); show(f"lookahead vta_max_chunks={mc}")
h.set_option("vta_max_chunks", 0)
# determinism of the look-ahead schedule
D.fill_uniform_(A, 0); D.householder_(A, al, 0); torch.cuda.synchronize(); A1 = A.clone()
D.fill_uniform_(A, 0); D.householder_(A, al, 0); torch.cuda.synchronize()
print("   lookahead bitwise repeatable:", bool(torch.equal(A, A1)), flush=True)
h.set_option("cvy_stagger", 0); show("lookahead, cvy_stagger=0")
h.set_option("cvy_stagger", 1); show("lookahead, cvy_stagger=1")
h.set_option("lookahead", 0)
h.set_option("cvy_stagger", 0); show("serial, cvy_stagger=0")
h.set_option("cvy_stagger", 1); show("serial, cvy_stagger=1")
h.set_option("lookahead", 1)
for lv, pc in ((1, 64), (1, 48), (1, 40), (2, 48), (2, 40)):
    h.set_option("panel_levels", lv); h.set_option("panel_ctas", pc); show(f"lookahead panel_levels={lv} panel_ctas={pc}")
print("   resid:", resid(), flush=True)
h.set_option("panel_levels", 2); h.set_option("panel_ctas", 0)
