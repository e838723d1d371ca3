"""Multi-GPU (NCCL) parity: world_size 2 on one box, DArray-style column blocks; skipped when < 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, m, n, nb, out):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import dhqr_b200 as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        h = D.init_distributed(rank)
        b = D.splits(world, n)
        c0, nl = b[rank], b[rank + 1] - b[rank]
        Al = D.colmajor_empty(m, nl, f"cuda:{rank}")
        D.fill_uniform_(Al, 0, 0, c0, h)
        Ad = D.ColumnBlockMatrix(Al, n, c0, h)
        H = D.qr_(Ad, nb=nb)
        rhs = D.colmajor_empty(m, 1, f"cuda:{rank}")
        D.fill_uniform_(rhs, 1, 0, 0, h)
        x = D.ldiv(H, rhs[:, 0].contiguous())
        torch.cuda.synchronize()
        gathered = [None] * world
        dist.all_gather_object(gathered, (c0, Al.cpu().numpy(), H.α.cpu().numpy(), x.cpu().numpy()))
        if rank == 0:
            g = sorted(gathered, key=lambda t: t[0])
            np.savez(out, H=np.hstack([t[1] for t in g]), alphas=np.stack([t[2] for t in g]), xs=np.stack([t[3] for t in g]))
        D.shutdown_distributed()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [(2048, 512, 0), (1500, 333, 0), (1024, 128, 1)])
def test_world2_nccl_matches_oracle(tmp_path, case, oracle, coracle):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    m, n, nb = case
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), m, n, nb, out), nprocs=2, join=True)
    g = np.load(out)
    A0 = coracle.fill_uniform(0, m, n)
    b = oracle.np_uniform(1, m, 1)[:, 0].copy()
    Href = A0.copy(order="F")
    Href, aref = coracle.qr(Href)
    xr = coracle.ldiv(Href, aref, b)
    assert np.abs(g["H"] - Href).max() < 1e-10
    for r in range(2):                                               # alpha and x replicated on every rank
        assert np.abs(g["alphas"][r] - aref).max() < 1e-12 * np.abs(aref).max()
        assert np.abs(g["xs"][r] - xr).max() < 1e-9 * np.abs(xr).max()
