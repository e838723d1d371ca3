"""Host-side mirror of the reference interface: layout and partition logic (no GPU needed)."""
import numpy as np
import pytest
import torch

import dhqr_b200 as D


def test_splits_match_darray_default_distribution():
    # DistributedArrays defaultdist: even chunks, remainder to the first blocks (T:71 uses (1, nworkers()))
    assert D.splits(2, 100) == [0, 50, 100]
    assert D.splits(4, 4096) == [0, 1024, 2048, 3072, 4096]
    assert D.splits(3, 10) == [0, 4, 7, 10]
    assert D.splits(8, 8192)[-1] == 8192 and len(D.splits(8, 8192)) == 9
    for P in range(1, 9):
        b = D.splits(P, 103)
        assert b[0] == 0 and b[-1] == 103 and all(0 <= b[i + 1] - b[i] <= 103 // P + 1 for i in range(P))


def test_colmajor_helpers_cpu():
    A = D.colmajor_empty(5, 3, device="cpu")
    assert A.shape == (5, 3) and A.stride() == (1, 5)
    B = D.colmajor_empty(5, 3, device="cpu", lda=8)
    assert B.stride() == (1, 8)
    x = np.arange(12.0).reshape(4, 3)
    C = D.to_colmajor(x, device="cpu")
    assert C.stride() == (1, 4) and np.array_equal(C.numpy(), x)


def test_local_column_block_indexing():
    # LocalColumnBlock (S:26-40): global column j lives at local column j - dj
    Al = D.to_colmajor(np.arange(20.0).reshape(4, 5), device="cpu")
    blk = D.LocalColumnBlock(Al, 10, range(10, 15))
    assert torch.equal(blk.global_col(12), Al[:, 2])


def test_alphafactor():
    assert D.alphafactor(3.0) == -1.0 and D.alphafactor(-2.0) == 1.0 and D.alphafactor(0.0) == 0.0


def test_rejects_row_major_and_wrong_dtype():
    from dhqr_b200.api import _lda
    with pytest.raises(ValueError):
        _lda(torch.zeros(4, 3, dtype=torch.float64))            # row-major
    with pytest.raises(TypeError):
        _lda(D.colmajor_empty(4, 3, device="cpu").float())


def test_bench_column_norm_property_on_oracle_output():
    """bench.py's rank-local parity property (||R[0:j+1, j]|| == ||A0[:, j]||) holds for the oracle's factorisation of any
    column block and trips on a corrupted one."""
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "oracle"))
    import bench, dhqr_oracle as O
    co = O.COracle()
    m, n = 200, 64
    A0 = co.fill_uniform(3, m, n)
    H, alpha = co.qr(A0.copy(order="F"))
    for c0, nl in ((0, 64), (0, 24), (24, 40)):
        blk = lambda X: torch.from_numpy(np.ascontiguousarray(X[:, c0:c0 + nl]))
        assert bench.column_norm_defect(torch, blk(H), torch.from_numpy(alpha), blk(A0), n, c0) < 1e-13
    Hbad = H.copy(); Hbad[3, 40] += 0.5
    assert bench.column_norm_defect(torch, torch.from_numpy(np.ascontiguousarray(Hbad[:, 24:])), torch.from_numpy(alpha),
                                    torch.from_numpy(np.ascontiguousarray(A0[:, 24:])), n, 24) > 1e-3


def test_tools_and_entry_scripts_compile():
    """Every helper script shipped in the repo at least parses (they only run on a GPU box)."""
    import glob, os, py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "tools", "*.py")) + [os.path.join(root, f) for f in ("bench.py", "__graft_entry__.py", "dhqr_b200.py")]
    assert len(files) > 5
    for f in files:
        py_compile.compile(f, doraise=True)


def test_balanced_splits_follow_the_reference_formulas():
    # T:35: round((N / sqrt(np)) * sqrt(p))   and   T:36: round(N * (1 - sqrt((np - p) / np)))
    for P, n in ((2, 4096), (4, 4096), (8, 8192), (3, 100), (4, 3)):
        for rule, f in (("trailing", lambda p: n * (p / P) ** 0.5), ("upstream", lambda p: n * (1 - ((P - p) / P) ** 0.5))):
            b = D.balanced_splits(P, n, rule)
            assert len(b) == P + 1 and b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(P))
            for p in range(1, P):
                assert b[p] == max(b[p - 1], int(round(f(p))))
    assert D.balanced_splits(2, 4096, "upstream") == [0, 1200, 4096]
    assert D.balanced_splits(2, 4096) == [0, 2896, 4096]
    with pytest.raises(ValueError):
        D.balanced_splits(2, 10, "nope")
    # trailing-update work of the right-looking factorisation per rank: reflector j (length m - j) touches the local columns
    # right of j.  The T:35 split evens it out compared with the default distribution; the T:36 split does the opposite.
    m, n, P = 32768, 4096, 4

    def work(bounds):
        j = np.arange(n)
        return np.array([float(np.sum((m - j) * np.clip(bounds[p + 1] - np.maximum(j + 1, bounds[p]), 0, None))) for p in range(P)])

    imb = lambda w: w.max() / w.mean()
    assert imb(work(D.balanced_splits(P, n))) < 1.1 < imb(work(D.splits(P, n))) < imb(work(D.balanced_splits(P, n, "upstream")))
