"""Warm Q'b at BASELINE config 3 (one right-hand side): the GEMV sweep with T' of every panel computed first (option qt_vec = 1)
against the GEMM-shaped block update per panel (qt_vec = 0); same b, results compared; per-kernel-class profile of the new path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dhqr_b200 as D
dev = torch.device("cuda:0"); h = D.default_handle(0)
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 4096)
A = D.colmajor_empty(m, n, dev); D.fill_uniform_(A, 0)
H = D.qr_(A)
b = torch.rand(m, dtype=torch.float64, device=dev)
out = {}
for vec in (1, 0, 1):
    h.set_option("qt_vec", vec)
    ts = []
    for it in range(6):
        w = b.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.apply_qt_(w, A, h); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out[vec] = w
    print(f"qt_vec={vec}: " + " ".join(f"{t:.3f}" for t in ts) + " ms", flush=True)
print("rel diff vec vs gemm path:", float((out[1] - out[0]).norm() / out[0].norm()), " |Q'b|/|b| - 1:", float(out[1].norm() / b.norm() - 1))
w = out[1].clone(); D.apply_q_(w, A, h)
print("|Q Q'b - b|/|b|:", float((w - b).norm() / b.norm()))
h.set_option("qt_vec", 1); h.set_option("profile", 1); h.profile_reset()
w = b.clone(); D.apply_qt_(w, A, h); torch.cuda.synchronize()
for name, r in h.profile().items():
    if r["count"]:
        print(f"  {name:16s} {r['ms']:8.3f} ms {r['count']:5d} launches")
h.set_option("profile", 0)
