"""dhqr_qr_host_f64 at BASELINE config 3 from pinned host memory: wall time per upload plan (chunk width, assumed link speed),
one stage trace, and cuBLAS DGEMM at the shapes of the two bulk GEMMs (what the library reaches at K = 128)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import dhqr_b200 as D
dev = torch.device("cuda:0"); h = D.default_handle(0)
m, n = 32768, 4096
host = torch.empty((n, m), dtype=torch.float64).pin_memory()
src = torch.empty((n, m), dtype=torch.float64, device=dev)
D.fill_uniform_(src.t(), 0)
al = torch.empty(n, dtype=torch.float64).pin_memory()
def run(reps=4):
    ts = []
    for _ in range(reps):
        host.copy_(src); torch.cuda.synchronize()
        t0 = time.perf_counter()
        D._lib.call("dhqr_qr_host_f64", h.raw, m, n, C.c_void_p(host.data_ptr()), m, C.c_void_p(al.data_ptr()), 0)
        ts.append((time.perf_counter() - t0) * 1e3)
    return ts
run(1)
PLANS = [(512, 50, 27)] if "quick" in sys.argv else [(512, 50, 27), (0, 50, 27), (256, 50, 27), (1024, 50, 27), (512, 40, 27), (512, 56, 24), (384, 50, 27), (768, 50, 27)]
for chunk, gbs, tf in PLANS:
    h.set_option("host_chunk", chunk); h.set_option("host_h2d_gbs", gbs); h.set_option("host_tflops", tf)
    ts = run()
    print(f"chunk {chunk:5d} link {gbs:3d} GB/s dev {tf} TF: " + " ".join(f"{t:.2f}" for t in ts) + " ms", flush=True)
h.set_option("host_chunk", 512); h.set_option("host_h2d_gbs", 50); h.set_option("host_tflops", 27)
h.set_option("host_trace", 1)
run(1)
h.set_option("host_trace", 0)
# residual of the last run against the input
A = host.t().to(dev); A0 = src.t()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
if "quick" in sys.argv: sys.exit(0)
# cuBLAS at the bulk shapes
def tm(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
rows, nc = 32768, 3712
V = torch.rand(128, rows, dtype=torch.float64, device=dev).t()          # column-major rows x 128
Y = torch.rand(nc, 128, dtype=torch.float64, device=dev).t()            # column-major 128 x nc
Cm = torch.rand(nc, rows, dtype=torch.float64, device=dev)              # C' row-major = C column-major
t = tm(lambda: torch.addmm(Cm, Y.t(), V.t(), out=Cm))                    # C' += Y' V'  (same GEMM, K = 128)
print(f"cuBLAS dgemm C += V Y   ({rows}x{nc}, K=128): {t:.3f} ms = {2*rows*nc*128/t/1e9:.1f} TFLOP/s")
t = tm(lambda: torch.mm(Y.t()[:, :128].contiguous().t() if False else V.t(), Cm.t()))   # W = V' C (128 x nc, K = rows)
print(f"cuBLAS dgemm W  = V' C  (128x{nc}, K={rows}): {t:.3f} ms = {2*rows*nc*128/t/1e9:.1f} TFLOP/s")
a = torch.rand(8192, 8192, dtype=torch.float64, device=dev); b = torch.rand(8192, 8192, dtype=torch.float64, device=dev)
t = tm(lambda: torch.mm(a, b), 3)
print(f"cuBLAS dgemm 8192^3: {t:.3f} ms = {2*8192**3/t/1e9:.1f} TFLOP/s")
