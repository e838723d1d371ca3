"""numpy restatement of k_panel's fast path (distributedhouseholderqr.jl_b200/csrc/dhqr_kernels.cuh), stage by stage:
CholeskyQR2 (two Gram / Cholesky / triangular-solve passes with the kernel's guards) followed by Householder
reconstruction (LU of the top block of E - Q S with the signs picked on the fly, row-local solve below it), producing the
reference's storage (S:127-135: v scaled to |v|^2 = 2 in the lower trapezoid including the diagonal, R above, diag(R) in alpha).

Test infrastructure only (tests/test_fastpath_model.py): it pins on the CPU that the algorithm the kernel runs yields the
reflectors of the reference's column recurrences, and that the guards send ill-conditioned panels to the column path.
The triangular solves are the blocked form the kernel uses on the fp64 tensor pipe: 8-column blocks, explicitly inverted 8x8
diagonal blocks, products R_ab inv(R_bb) formed once.
"""
import numpy as np

IB = 32
SPREAD_MIN = 4e-3     # FAST_SPREAD_MIN of the kernel for the tensor-pipe solves (1e-5 with row-by-row substitution)


def blocked_trsm(X, R, dgi):
    """X <- X R^{-1} for upper-triangular R (IB x IB), dgi = 1 / diag(R); panel_trsm / trsm_dmma of the kernel."""
    X = np.array(X, dtype=np.float64, copy=True)
    nb = IB // 8
    Dv = []
    for b in range(nb):                                   # inverses of the diagonal blocks, one column per thread
        Rb = R[8 * b:8 * b + 8, 8 * b:8 * b + 8]
        inv = np.zeros((8, 8))
        for c in range(8):
            for i in range(7, -1, -1):
                s = 1.0 if i == c else 0.0
                for j in range(i + 1, 8):
                    s -= Rb[i, j] * inv[j, c]
                inv[i, c] = s * dgi[8 * b + i] if i <= c else 0.0
        Dv.append(inv)
    Wn = {(a, b): -(R[8 * a:8 * a + 8, 8 * b:8 * b + 8] @ Dv[b]) for b in range(nb) for a in range(b)}
    out = np.zeros_like(X)
    for b in range(nb):
        acc = X[:, 8 * b:8 * b + 8] @ Dv[b]
        for a in range(b):
            acc = acc + out[:, 8 * a:8 * a + 8] @ Wn[(a, b)]
        out[:, 8 * b:8 * b + 8] = acc
    return out


def cholesky_upper(G):
    """Right-looking upper Cholesky as in the kernel: returns (R, rinv, ok)."""
    g = np.array(G, dtype=np.float64, copy=True)
    R = np.zeros((IB, IB))
    rinv = np.zeros(IB)
    ok = True
    for j in range(IB):
        d = g[j, j]
        if not (d > 0.0) or not (d < 1e300):
            ok = False
            d = abs(d) + 1.0                                   # keep going with finite numbers; the result is discarded
        ri = 1.0 / np.sqrt(d)
        rinv[j] = ri
        R[j, j:] = g[j, j:] * ri
        R[j, j] = d * ri
        for i in range(j + 1, IB):
            g[i, i:] -= R[j, i] * R[j, i:]
    return R, rinv, ok


def fast_panel(P):
    """Returns (H, alpha, took_fast_path).  H, alpha are None when the guards ask for the column-by-column path."""
    P = np.array(P, dtype=np.float64)
    m, n = P.shape
    assert n == IB and m >= 2 * IB
    # pass 1
    R1, rinv1, ok = cholesky_upper(P.T @ P)
    d1 = np.diag(R1)
    if not ok or not (d1.min() > SPREAD_MIN * d1.max()):
        return None, None, False
    Q1 = blocked_trsm(P, R1, rinv1)
    # pass 2 with the orthogonality guard of the kernel
    G2 = Q1.T @ Q1
    if not np.all(np.abs(G2 - np.eye(IB)) <= 0.25 / IB):
        return None, None, False
    R2, rinv2, ok = cholesky_upper(G2)
    if not ok:
        return None, None, False
    Q = blocked_trsm(Q1, R2, rinv2)
    Rt = np.triu(R2 @ R1)
    # Householder reconstruction: LU of the top block of E - Q S, signs on the fly (CTA 0 of the kernel)
    Wt = Q[:IB, :].copy()
    Sg = np.zeros(IB)
    Ud = np.zeros(IB)
    for j in range(IB):
        w = Wt[j, j]
        Sg[j] = -1.0 if w > 0.0 else 1.0
        Ud[j] = 1.0 + abs(w)
        f = Sg[j] / Ud[j]
        Wt[j + 1:, j + 1:] += np.outer(f * Wt[j + 1:, j], Wt[j, j + 1:])
    rsq = 1.0 / np.sqrt(Ud)
    cl = -Sg / Ud
    # rows below the top block: V = M Rr^{-1}, Rr = diag(sqrt(Ud)) (I + diag(cl) striu(U))
    U = np.triu(Wt, 1)
    sq = Ud * rsq
    Rr = np.diag(sq) + (cl * sq)[:, None] * U
    H = np.zeros((m, IB))
    H[IB:, :] = blocked_trsm(Q[IB:, :], Rr, rsq)
    # top block: V below the diagonal, R above, alpha
    for i in range(IB):
        for j in range(IB):
            if i > j:
                H[i, j] = Wt[i, j] * rsq[j]
            elif i == j:
                H[i, j] = -Sg[j] * (Ud[j] * rsq[j])
            else:
                H[i, j] = Sg[i] * Rt[i, j]
    alpha = Sg * np.diag(Rt)
    return H, alpha, True
