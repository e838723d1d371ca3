#!/usr/bin/env python
"""bench.py — QR GFLOP/s (fp64) of qr! on the BASELINE workload, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--m M --n N --nb NB]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json): QR of a 32768 x 4096 fp64 matrix, A[i,j] ~ U[0,1) from the counter-based generator
(synthetic; mirrors rand at test/runtests.jl:45-46), DArray-style contiguous column blocks over the N GPUs
(strong scaling: total work fixed).  A "step" is one full factorisation qr!(A) of a fresh matrix.
  value : (2mn^2 - 2/3 n^3) / t, inputs resident in HBM, CUDA events, max over ranks.
  e2e   : the same through the host-buffer entry (pinned host A -> H2D -> factor -> D2H of A and alpha).
  roofline : the dominant kernel class (gemm_cvy, C += V*Y on the fp64 tensor pipe), algorithmic flops /
             CUDA-event time of that class, against a cuBLAS DGEMM burst measured in this run
             (MEASURED_PEAKS.json carries no fp64 entry).
  cpu_baseline : the oracle's C restatement of the reference algorithm on the host cores, bounded sample.
--impl reference times that CPU restatement alone (the reference is Julia; Julia is not installed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "QR GFLOP/s (fp64)"


def qr_flops(m, n):
    return 2.0 * m * n * n - 2.0 / 3.0 * n ** 3


# ---------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()                      # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement on a bounded sample (first J column steps of S:127 on the full matrix)
# ---------------------------------------------------------------------------------------------
def cpu_sample(m, n, target_s=12.0, jstop=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dhqr_oracle as O
    co = O.COracle()
    cores = int(os.environ.get("DHQR_CPU_THREADS", "0")) or co.max_threads()
    A = co.fill_uniform(0, m, n)
    if jstop is None:
        best = None
        for nt in sorted({cores, max(1, cores // 2), max(1, cores // 4)}, reverse=True):   # calibration: 32 column steps per
            co.qr_steps(A, 2, nt)                                       # thread count (2 untimed steps first: thread start-up,
            A = co.fill_uniform(0, m, n)                                # page touch); all / half / a quarter of the host threads
            t = time.perf_counter()
            co.qr_steps(A, 32, nt)
            dt = max(time.perf_counter() - t, 1e-4) / 2.0
            if best is None or dt < best[0]:
                best = (dt, nt)
            A = co.fill_uniform(0, m, n)
        dt, cores = best
        os.environ["DHQR_CPU_THREADS"] = str(cores)
        jstop = int(max(8, min(n, 16 * target_s / dt)))
    t = time.perf_counter()
    _, fl = co.qr_steps(A, jstop, cores)
    dt = time.perf_counter() - t
    return {"value": fl / dt / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "sample": f"first {jstop} of {n} column steps (S:127-144) of the {m}x{n} factorisation, {dt:.2f} s, "
                      f"OpenMP over trailing-column chunks as S:203-211", "seconds": dt, "jstop": jstop}


def run_reference(args):
    """--impl reference: the reference's CPU path.  Julia is absent, so this is the oracle port (kind=port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m, n = args.m, args.n
    vals, jstop = [], None
    for it in range(args.warmup + args.steps):
        s = cpu_sample(m, n, target_s=max(2.0, 40.0 / max(1, args.warmup + args.steps)), jstop=jstop)
        jstop = s["jstop"]
        if it >= args.warmup:
            vals.append(s)
    tot_s = sum(v["seconds"] for v in vals)
    val = sum(v["value"] * v["seconds"] for v in vals) / tot_s
    cb = {"value": val, "unit": "GFLOP/s", "cores": vals[0]["cores"], "kind": "port", "sample": vals[0]["sample"]}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / len(vals),
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"qr! of {m}x{n} fp64 (bounded sample per step)", "m": m, "n": n},
                      "cpu_baseline": cb,
                      "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dhqr_b200 as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        h = D.init_distributed(local)
    else:
        h = D.default_handle(local)

    m, n, nb = args.m, args.n, args.nb
    b = D.splits(world, n)
    c0, nl = b[rank], b[rank + 1] - b[rank]
    flops = qr_flops(m, n)
    K, W = args.steps, args.warmup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxover(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumover(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # pool of fresh matrices so the timed region holds only qr! calls (inputs resident in HBM)
    pool_n = max(1, min(K, args.pool))
    pool = [D.colmajor_empty(m, nl, dev) for _ in range(pool_n)]
    alpha = torch.zeros(n, dtype=torch.float64, device=dev)

    def refill():
        for A in pool:
            D.fill_uniform_(A, 0, 0, c0, h)

    def step(A):
        D.householder_(D.ColumnBlockMatrix(A, n, c0, h) if world > 1 else A, alpha, nb, h)

    refill()
    for w in range(W):
        step(pool[w % pool_n])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, done, l0 = 0.0, 0, h.launch_count()
    launches = 0
    while done < K:
        g = min(pool_n, K - done)
        refill()
        barrier()
        la = h.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(g):
            step(pool[i])
        e1.record()
        barrier()
        total_ms += maxover(e0.elapsed_time(e1))
        launches += h.launch_count() - la
        done += g
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / K
    value = flops / (ms_per_step * 1e-3) / 1e9
    launches_all = int(sumover(float(launches)))

    # parity of the last timed factorisation: ||QR - A||_F / ||A||_F (single GPU; multi-GPU checked in tests)
    resid, colnorm = None, None
    if world == 1 and not args.no_check:
        resid = gpu_residual(torch, D, pool[(K - 1) % pool_n], alpha, m, n, dev)
    if not args.no_check:
        # size-independent and rank-local: Q orthogonal => ||A0[:, j]|| == ||R[0:j+1, j]|| for every column of the block
        try:
            A0 = D.colmajor_empty(m, nl, dev)
            D.fill_uniform_(A0, 0, 0, c0, h)
            local_defect = column_norm_defect(torch, pool[(K - 1) % pool_n], alpha, A0, n, c0)
            del A0
        except Exception as e:   # a failing check must not take the timing line (or the other ranks) with it
            sys.stderr.write(f"[bench] column-norm check failed on rank {rank}: {type(e).__name__}: {e}\n")
            local_defect = float("inf")
        colnorm = maxover(local_defect)   # the collective is outside the try: every rank reaches it
        if colnorm != colnorm or colnorm in (float("inf"), float("-inf")):
            colnorm = "check failed (see stderr)"   # keep the line strict JSON

    # ---- per-kernel-class profile (separate, untimed step) -> roofline of the dominant kernel ----
    h.set_option("profile", 1)
    refill()
    barrier()
    h.profile_reset()
    step(pool[0])
    torch.cuda.synchronize()
    prof = h.profile()
    h.set_option("profile", 0)
    roof = None
    if world == 1 or rank == 0:
        peak = dgemm_peak(torch, dev)
        dom = max((k for k in prof if k.startswith("k_gemm")), key=lambda k: prof[k]["ms"], default=None)
        tot = sum(v["ms"] for v in prof.values())
        if dom:
            ach = prof[dom]["work"] / (prof[dom]["ms"] * 1e-3) / 1e12
            traffic, tnote = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[dom]
                traffic, tnote = tj["dram_bytes_per_launch"], f"dram bytes of the largest launch ({tj['captured_launch']}), ncu --set full, {tj['source']}"
            except Exception:
                pass
            roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_note": tnote, "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                    "launches": prof[dom]["count"], "avg_launch_ms": prof[dom]["ms"] / max(1, prof[dom]["count"]),
                    "share_of_step": prof[dom]["ms"] / tot if tot else None,
                    "classes": {k: {"ms": round(v["ms"], 3), "count": v["count"],
                                    "tflops": (v["work"] / (v["ms"] * 1e-3) / 1e12 if k.startswith("k_gemm") and v["ms"] > 0 else None)}
                                for k, v in prof.items()}}

    # ---- e2e: host buffers through the reference-facing entry ----
    e2e = None
    if not args.no_e2e:
        hostA = torch.empty((nl, m), dtype=torch.float64).pin_memory().t()      # column-major pinned (m, nl)
        src = D.colmajor_empty(m, nl, dev)
        D.fill_uniform_(src, 0, 0, c0, h)
        hostA.copy_(src)
        pristine = hostA.clone()
        host_alpha = torch.empty(n, dtype=torch.float64).pin_memory()
        Ke = max(1, min(K, args.e2e_steps))
        tot = 0.0
        for it in range(1 + Ke):
            hostA.copy_(pristine)
            barrier()
            t0 = time.perf_counter()
            if world == 1:
                import ctypes as C
                D._lib.call("dhqr_qr_host_f64", h.raw, m, n, C.c_void_p(hostA.data_ptr()), m, C.c_void_p(host_alpha.data_ptr()), nb)
            else:
                dA = pool[0]
                dA.copy_(hostA, non_blocking=True)
                step(dA)
                hostA.copy_(dA, non_blocking=True)
                host_alpha.copy_(alpha, non_blocking=True)
                torch.cuda.synchronize()
            dt = maxover(time.perf_counter() - t0)
            if it > 0:
                tot += dt
        e2e = {"value": flops / (tot / Ke) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(sumover(float(m * nl * 8))),
               "d2h_bytes_per_step": int(sumover(float(m * nl * 8))) + n * 8, "ms_per_step": 1e3 * tot / Ke, "steps": Ke,
               "path": "dhqr_qr_host_f64 (C-ABI, pinned host buffers)" if world == 1 else "pinned host block -> qr_ -> host (per rank)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_sample(m, n, target_s=args.cpu_seconds)
        cpu.pop("seconds", None), cpu.pop("jstop", None)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": {"workload": f"qr! of {m}x{n} fp64, DArray-style contiguous column blocks over {world} GPU(s)",
                          "m": m, "n": n, "nb": nb or 128, "cols_per_gpu": nl, "l2": "inputs (m*n*8 B per step) larger than L2; fresh matrix per step",
                          "timing": "CUDA events around K back-to-back qr! calls on fresh matrices, max over ranks"},
               "clocks": clocks, "gpu_launches": launches_all, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu,
               "parity": {"qr_residual_fro_rel": resid, "tolerance": 1e-13,
                          "column_norm_defect_max_rel": colnorm, "column_norm_tolerance": 1e-12}}
        print(json.dumps(out))
    if world > 1:
        D.shutdown_distributed()
        dist.destroy_process_group()


def column_norm_defect(torch, H, alpha, A0, n, c0):
    """max_j | ||R[0:j+1, j]|| / ||A0[:, j]|| - 1 | over the columns of one column block (H = factored block, global
    columns c0..; R's strict upper part sits above the global diagonal, diag(R) in alpha)."""
    nl = H.shape[1]
    U = torch.triu(H[:n], diagonal=1 - c0)          # keeps (i, j) with i < c0 + j
    r2 = (U * U).sum(0) + alpha[c0:c0 + nl] ** 2
    a2 = (A0 * A0).sum(0)
    return float((torch.sqrt(r2 / a2) - 1.0).abs().max().item())


def gpu_residual(torch, D, A, alpha, m, n, dev):
    """||QR - A0||_F / ||A0||_F with Q applied as block reflectors (torch fp64 on the GPU; size independent)."""
    R = torch.zeros(m, n, dtype=torch.float64, device=dev)
    R[:n] = torch.triu(A[:n], 1) + torch.diag(alpha)
    nbk = 128
    for k in range(((n - 1) // nbk) * nbk, -1, -nbk):
        kb = min(nbk, n - k)
        V = torch.tril(A[k:, k:k + kb])
        Tinv = torch.eye(kb, dtype=torch.float64, device=dev) + torch.triu(V.T @ V, 1)
        R[k:] -= V @ torch.linalg.solve_triangular(Tinv, V.T @ R[k:], upper=True)
    A0 = D.colmajor_empty(m, n, dev)
    D.fill_uniform_(A0, 0)
    return float(torch.linalg.norm(R - A0) / torch.linalg.norm(A0))


def dgemm_peak(torch, dev, nn=8192):
    a = torch.rand(nn, nn, dtype=torch.float64, device=dev)
    b = torch.rand(nn, nn, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * nn ** 3 / (best * 1e-3) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=32768)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
