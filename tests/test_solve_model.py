"""CPU restatements of the second session's device algorithms, pinned on the oracle:
 * Q'b with one right-hand side: T' of every 128-column panel from its Gram matrix, then per panel w = V'b, y = -T'w, b += V y with V
   read in place (lower trapezoid including the diagonal) - must equal the reference's reflector-by-reflector sweep (S:232-242);
   Qb is the reverse sweep with T instead of T'.
 * the symmetric Gram kernel: 64-row chunks, the 10 blocks of 32 x 32 on or above the diagonal (two 32 x 16 halves each), mirrored
   writes, split-K partials summed four lanes per element - must equal V'V."""
import numpy as np
import pytest


def qt_sweep(h, b, nb=128, trans=False):
    m, n = h.shape
    w = np.array(b, dtype=np.float64, copy=True)
    starts = list(range(0, n, nb))
    T = {}
    for c0 in starts:                                    # prepare: independent of b
        kb = min(nb, n - c0)
        V = np.tril(h[c0:, c0:c0 + kb])
        G = V.T @ V
        T[c0] = np.linalg.inv(np.eye(kb) + np.tril(G, -1))          # T' = (I + stril(V'V))^-1   (|v|^2 = 2)
    for c0 in (reversed(starts) if trans else starts):   # sweep: sequential in b
        kb = min(nb, n - c0)
        V = np.tril(h[c0:, c0:c0 + kb])
        y = -(T[c0].T if trans else T[c0]) @ (V.T @ w[c0:])
        w[c0:] += V @ y
    return w


@pytest.mark.parametrize("mn", [(300, 40), (700, 300), (1001, 337), (1500, 1024)])
def test_vector_sweep_equals_the_reference_sweep(oracle, mn):
    m, n = mn
    A = oracle.np_uniform(31, m, n)
    H, alpha = oracle.np_qr(A)
    b = oracle.np_uniform(32, m, 1)[:, 0]
    ref = oracle.np_apply_qt(H, b)
    got = qt_sweep(H, b)
    assert np.linalg.norm(got - ref) < 1e-13 * np.linalg.norm(b)
    back = qt_sweep(H, got, trans=True)                  # Q (Q'b) = b
    assert np.linalg.norm(back - b) < 1e-13 * np.linalg.norm(b)


def gram_sym(V, nsplit):
    rows = V.shape[0]
    nchunks = (rows + 63) // 64
    Vp = np.zeros((nchunks * 64, 128))
    Vp[:rows] = V
    cps = (nchunks + nsplit - 1) // nsplit
    parts = np.zeros((nsplit, 128, 128))
    blocks = [(bi, bj) for bi in range(4) for bj in range(bi, 4)]
    assert len(blocks) == 10
    for s in range(nsplit):
        for ch in range(s * cps, min(nchunks, (s + 1) * cps)):
            Vc = Vp[ch * 64:(ch + 1) * 64]
            for bi, bj in blocks:
                for h in range(2):                       # the two warps of a block: 32 x 16 halves
                    r = slice(bi * 32, bi * 32 + 32)
                    c = slice(bj * 32 + h * 16, bj * 32 + h * 16 + 16)
                    g = Vc[:, r].T @ Vc[:, c]
                    parts[s][r, c] += g
                    if bi != bj:
                        parts[s][c, r] += g.T
    lanes = [sum(parts[p] for p in range(q, nsplit, 4)) if q < nsplit else 0.0 for q in range(4)]   # k_wreduce4: lane q sums p = q, q+4, ...
    return (lanes[0] + lanes[1]) + (lanes[2] + lanes[3])


@pytest.mark.parametrize("rows,nsplit", [(128, 1), (1000, 4), (4097, 13), (8192, 128)])
def test_symmetric_gram_blocks_cover_the_matrix(oracle, rows, nsplit):
    V = oracle.np_uniform(33, rows, 128)
    G = gram_sym(V, nsplit)
    ref = V.T @ V
    assert np.abs(G - ref).max() < 1e-12 * np.abs(ref).max()
    assert np.array_equal(G, G.T)
