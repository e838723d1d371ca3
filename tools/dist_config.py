"""BASELINE configs 4/5 under torchrun: DArray-style column blocks over N GPUs, full qr! + H \\ b, checked through
size-independent properties on rank 0 (||QR - A||/||A||, ||Q'b|| = ||b||, normal equations vs cuSOLVER lstsq).
    torchrun --nproc-per-node 8 tools/dist_config.py 65536 8192"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import dhqr_b200 as D
m, n = int(sys.argv[1]), int(sys.argv[2])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
h = D.init_distributed(local)
b = D.splits(world, n); c0, nl = b[rank], b[rank + 1] - b[rank]
Al = D.colmajor_empty(m, nl, dev); D.fill_uniform_(Al, 0, 0, c0, h)
Ad = D.ColumnBlockMatrix(Al, n, c0, h)
torch.cuda.synchronize(); dist.barrier()
ts = []
for rep in range(3):
    D.fill_uniform_(Al, 0, 0, c0, h); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); H = D.qr_(Ad); torch.cuda.synchronize(); dist.barrier(); ts.append(time.perf_counter() - t0)
rhs = D.colmajor_empty(m, 1, dev); D.fill_uniform_(rhs, 1, 0, 0, h)
bvec = rhs[:, 0].contiguous()
t0 = time.perf_counter(); x = D.ldiv(H, bvec); torch.cuda.synchronize(); tcold = time.perf_counter() - t0   # first call: workspace growth
tw = []
for rep in range(4):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); x = D.ldiv(H, bvec); torch.cuda.synchronize(); dist.barrier(); tw.append(time.perf_counter() - t0)
tsolve = min(tw)
work = bvec.clone(); tq = []; tb = []
for rep in range(3):
    work.copy_(bvec); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); D.apply_qt_(work, Ad); torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    D.backsolve_(work, Ad, H.α); torch.cuda.synchronize(); dist.barrier(); tq.append(t1 - t0); tb.append(time.perf_counter() - t1)
qtb = D.apply_qt_(bvec.clone(), Ad)
# gather the factored blocks on rank 0 (column blocks are contiguous: concatenate)
blocks = [torch.empty((b[r + 1] - b[r], m), dtype=torch.float64, device=dev) for r in range(world)] if rank == 0 else None
dist.gather(Al.t().contiguous(), blocks, dst=0)
if rank == 0:
    fl = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3
    print(f"{world} GPUs, {m}x{n}: qr! {min(ts)*1e3:.1f} ms = {fl/min(ts)/1e12:.2f} TFLOP/s aggregate; H \\ b warm {tsolve*1e3:.2f} ms (Q'b {min(tq)*1e3:.2f} + back-substitution {min(tb)*1e3:.2f}; first call {tcold*1e3:.0f} ms)", flush=True)
    Hf = torch.cat(blocks, 0).t()                         # (m, n) view, column-major
    A0 = D.colmajor_empty(m, n, dev); D.fill_uniform_(A0, 0, 0, 0, D.Handle(local)) if False else None
    hh = D.Handle(local); D.fill_uniform_(A0, 0, 0, 0, hh)
    R = torch.zeros(m, n, dtype=torch.float64, device=dev)
    R[:n] = torch.triu(Hf[:n], 1) + torch.diag(H.α)
    for k in range(((n - 1) // 128) * 128, -1, -128):
        kb = min(128, n - k); V = torch.tril(Hf[k:, k:k + kb])
        Tinv = torch.eye(kb, dtype=torch.float64, device=dev) + torch.triu(V.T @ V, 1)
        R[k:] -= V @ torch.linalg.solve_triangular(Tinv, V.T @ R[k:], upper=True)
    res = float(torch.linalg.norm(R - A0) / torch.linalg.norm(A0))
    qn = abs(float(torch.linalg.norm(qtb) / torch.linalg.norm(bvec)) - 1.0)
    r = A0.T @ (A0 @ x) - A0.T @ bvec
    print(f"   ||QR-A||/||A|| = {res:.3e} (tol 1e-13);  | ||Q'b||/||b|| - 1 | = {qn:.2e};  ||A'Ax - A'b|| = {float(torch.linalg.norm(r)):.3e}", flush=True)
    assert res < 1e-13 and qn < 1e-12
    print("   OK", flush=True)
D.shutdown_distributed(); dist.destroy_process_group()
