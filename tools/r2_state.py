"""Bench workload status: default vs wide_panel=0, serial per-class profile, look-ahead timeline."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import dhqr_b200 as D
dev = torch.device("cuda:0"); h = D.default_handle(0)
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 4096)
A = D.colmajor_empty(m, n, dev); al = torch.zeros(n, dtype=torch.float64, device=dev)
fl = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3
def timeit(reps=4):
    ts = []
    for _ in range(reps + 1):
        D.fill_uniform_(A, 0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.householder_(A, al, 0); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:]), float(np.median(ts[1:]))
def show(tag):
    t, md = timeit(); print(f"{tag}: min {t:.2f} ms median {md:.2f} ms  {fl / t / 1e9:.2f} TFLOP/s", flush=True)
def profile(tag):
    h.set_option("profile", 1); D.fill_uniform_(A, 0); torch.cuda.synchronize(); h.profile_reset()
    D.householder_(A, al, 0); torch.cuda.synchronize(); p = h.profile(); h.set_option("profile", 0)
    print(tag, "serial profile [ms, launches, TFLOP/s]:", json.dumps({k: (round(v["ms"], 3), v["count"], round(v["work"] / v["ms"] / 1e9, 1) if v["work"] > 0 and v["ms"] > 0 and not k.startswith("k_panel") else None) for k, v in p.items()}), "sum", round(sum(v["ms"] for v in p.values()), 2), flush=True)
def timeline(tag):
    K = n // 128
    for rep in range(2):
        D.fill_uniform_(A, 0); torch.cuda.synchronize(); h.set_option("la_trace", rep)
        D.householder_(A, al, 0); torch.cuda.synchronize()
    buf = torch.zeros(3 * K, dtype=torch.float64, device=dev)
    D._lib.call("dhqr_debug_copy_f64", h.raw, b"la_times", C.c_void_p(buf.data_ptr()), 3 * K, None)
    t = buf.cpu().numpy().reshape(K, 3); h.set_option("la_trace", 0)
    dp = np.diff(np.concatenate([[0], t[:, 0]])); db = np.diff(np.concatenate([[0], t[:, 2]]))
    print(tag, "timeline total %.2f ms; panel steps:" % t[-1].max(), np.round(dp, 2).tolist(), flush=True)
    print(tag, "          bulk steps:", np.round(db, 2).tolist(), flush=True)
def stamps(tag):
    h.set_option("wide_trace", 1); D.fill_uniform_(A, 0); D.householder_(A, al, 0); torch.cuda.synchronize(); h.set_option("wide_trace", 0)
    buf = torch.zeros(32, dtype=torch.float64, device=dev)
    D._lib.call("dhqr_debug_copy_f64", h.raw, b"wstamps", C.c_void_p(buf.data_ptr()), 32, None); torch.cuda.synchronize()
    s = buf.cpu().numpy().view(np.int64)
    print(tag, "clock64 stamps of the last panel (cycles): chol1 [load+guard, factor, inverse, outputs]", s[0:4].tolist(), "chol2", s[8:12].tolist(), "hr128 [load, LU, top block + Rr inverse]", s[20:23].tolist(), flush=True)
show("default (wide)")
stamps("wide")
print("wide panels / redone:", h.get_option("wide_panels"), h.get_option("wide_redone"))
profile("wide"); timeline("wide")
h.set_option("lookahead", 0); show("wide serial"); h.set_option("lookahead", 1)
h.set_option("wide_panel", 0); show("narrow (r01 chain)"); profile("narrow"); h.set_option("wide_panel", 1)
show("default again")
