"""dhqr_qr_host_f64 at BASELINE config 3: width of the first (exposed) upload x catch-up streams, deadline joins."""
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
for first in (0, 512, 768):
    for cus in (3, 2):
        for chunk in (512, 384):
            h.set_option("host_first", first); h.set_option("host_cu_streams", cus); h.set_option("host_chunk", chunk)
            run(1)
            ts = run()
            print(f"first {first:4d} chunk {chunk:4d} cu_streams {cus}: " + " ".join(f"{t:.2f}" for t in ts) + " ms", flush=True)
h.set_option("host_first", 0); h.set_option("host_cu_streams", 3); h.set_option("host_chunk", 512)
h.set_option("host_trace", 1)
run(1)
