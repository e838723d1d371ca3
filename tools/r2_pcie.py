"""Host link bandwidth as bench.py's e2e leg sees it (pinned torch tensors, 1 GiB, both directions, alone and concurrently),
and a stage timeline of dhqr_qr_host_f64."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import dhqr_b200 as D
dev = torch.device("cuda:0"); h = D.default_handle(0)
m, n = 32768, 4096
host = torch.empty((n, m), dtype=torch.float64).pin_memory()
host2 = torch.empty((n, m), dtype=torch.float64).pin_memory()
d1 = torch.empty((n, m), dtype=torch.float64, device=dev); d2 = torch.empty_like(d1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
gb = host.numel() * 8 / 1e9
print(f"H2D alone : {gb / t(lambda: d1.copy_(host, non_blocking=True)):.1f} GB/s")
print(f"D2H alone : {gb / t(lambda: host2.copy_(d2, non_blocking=True)):.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d1.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): host2.copy_(d2, non_blocking=True)
tb = t(both)
print(f"H2D + D2H concurrently: {tb * 1e3:.1f} ms for 1 GiB each way = {gb / tb:.1f} GB/s per direction")
A = host.t()
D.fill_uniform_(d1.t(), 0); host.copy_(d1)
al = torch.empty(n, dtype=torch.float64).pin_memory()
h.set_option("host_trace", 1)
for rep in range(3):
    host.copy_(d1); torch.cuda.synchronize()
    t0 = time.perf_counter()
    D._lib.call("dhqr_qr_host_f64", h.raw, m, n, C.c_void_p(host.data_ptr()), m, C.c_void_p(al.data_ptr()), 0)
    print(f"dhqr_qr_host_f64: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
