"""N > 1 on CPU: world_size-2 gloo run of the SPMD column-block algorithm (one exchange per panel),
written with the numpy twin, checked against the single-block oracle.  Covers the host-side partition
logic (splits / ColumnBlock descriptors / owner order) that the GPU path shares."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, m, n, nb, out):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import dhqr_b200 as D
    import dhqr_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = D.splits(world, n)
        c0, c1 = b[rank], b[rank + 1]
        Al = O.np_uniform(0, m, c1 - c0, 0, c0)                  # each rank generates only its own columns
        alpha = np.zeros(n)
        # SPMD sweep: owner factors a panel with the reference recurrences, broadcasts V + alpha slice (C2),
        # every rank applies the reflectors to its local columns right of the panel (S:198-213)
        for owner in range(world):
            for pc in range(b[owner], b[owner + 1], nb):
                kb = min(nb, b[owner + 1] - pc)
                if rank == owner:
                    lo = pc - c0
                    P, a = O.np_qr(Al[pc:, lo:lo + kb])
                    Al[pc:, lo:lo + kb] = P
                    alpha[pc:pc + kb] = a
                    V = np.tril(P)
                    buf = torch.from_numpy(np.ascontiguousarray(V))
                    abuf = torch.from_numpy(alpha[pc:pc + kb].copy())
                else:
                    buf = torch.empty((m - pc, kb), dtype=torch.float64)
                    abuf = torch.empty(kb, dtype=torch.float64)
                dist.broadcast(buf, src=owner)
                dist.broadcast(abuf, src=owner)
                alpha[pc:pc + kb] = abuf.numpy()
                V = buf.numpy()
                t0 = max(pc + kb, c0)
                if t0 < c1:
                    for j in range(kb):                           # reflector by reflector, rows pc+j:
                        v = V[j:, j]
                        blk = Al[pc + j:, t0 - c0:]
                        blk -= np.outer(v, v @ blk)
        gathered = [None] * world
        dist.all_gather_object(gathered, (c0, Al))
        if rank == 0:
            H = np.hstack([g[1] for g in sorted(gathered, key=lambda t: t[0])])
            np.savez(out, H=H, alpha=alpha)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mn", [(96, 40), (130, 33)])
def test_world2_column_blocks_match_oracle(tmp_path, mn, oracle):
    m, n = mn
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), m, n, 8, out), nprocs=2, join=True)
    g = np.load(out)
    Href, aref = oracle.np_qr(oracle.np_uniform(0, m, n))
    assert np.abs(g["H"] - Href).max() < 1e-12
    assert np.abs(g["alpha"] - aref).max() < 1e-12
