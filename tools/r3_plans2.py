"""dhqr_qr_host_f64 at BASELINE config 3: chunk width x catch-up streams with deadline joins (host_chain_us = 300)."""
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
def run(reps=3):
    ts = []
    for _ in range(reps):
        host.copy_(src); torch.cuda.synchronize()
        t0 = time.perf_counter()
        D._lib.call("dhqr_qr_host_f64", h.raw, m, n, C.c_void_p(host.data_ptr()), m, C.c_void_p(al.data_ptr()), 0)
        ts.append((time.perf_counter() - t0) * 1e3)
    return ts
run(1)
for chunk in (512, 640, 768, 1024):
    for cus in (2, 3):
        for chain in (300, 500):
            h.set_option("host_chunk", chunk); h.set_option("host_cu_streams", cus); h.set_option("host_chain_us", chain)
            run(1)
            ts = run()
            print(f"chunk {chunk:4d} cu_streams {cus} chain {chain:4d} us: " + " ".join(f"{t:.2f}" for t in ts) + " ms", flush=True)
