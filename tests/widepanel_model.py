"""numpy restatement of the 128-column panel chain (k_chol128 / k_vpk_rmul / k_hr128 / k_trimm128 in
distributedhouseholderqr.jl_b200/csrc/dhqr_kernels.cuh), stage by stage:

    G1 = P'P          R1 = chol(G1)     X1 = R1^{-1} (explicit, recursive doubling)      Q1 = P X1
    G2 = Q1'Q1        R2 = chol(G2)     X2 = R2^{-1}                                      (orthogonality guard on G2)
    Wt = Q1[:nb] X2   signed LU of Wt (Householder reconstruction: Ballard, Demmel, Grigori, Jacquelin, Nguyen, Solomonik 2014)
    Rr = diag(sqrt(Ud)) (I + diag(cl) striu(U))      X3 = X2 Rr^{-1}      V[nb:] = Q1[nb:] X3      Rt = R2 R1

producing the reference's storage (S:127-135: v scaled to |v|^2 = 2 in the lower trapezoid including the diagonal, R above,
diag(R) in alpha) for a whole outer panel with three grid-wide reductions instead of one per column.

Test infrastructure only (tests/test_widepanel_model.py): pins on the CPU that this algorithm yields the reflectors of the
reference's column recurrences and that the guards refuse ill-conditioned panels.
"""
import numpy as np

NB = 128
ORTH_MAX = 0.25          # guard on ||Q1'Q1 - I||: max-norm <= ORTH_MAX / nb  =>  2-norm <= 1/4
KAPPA_MAX = 1.0e3        # guard on ||D R1^{-1}||_F, D = diag(||p_j||): multiplying by the EXPLICIT inverse (a GEMM on the tensor
                         # pipe instead of a substitution) costs ~1e-17 x that number in ||QR - A|| / ||A|| (measured below)


def cholesky_upper(G):
    """Right-looking upper Cholesky; returns (R, ok)."""
    g = np.array(G, dtype=np.float64, copy=True)
    n = g.shape[0]
    R = np.zeros((n, n))
    ok = True
    for j in range(n):
        d = g[j, j]
        if not (d > 0.0) or not (d < 1e300):
            ok = False
            d = abs(d) + 1.0
        ri = 1.0 / np.sqrt(d)
        R[j, j:] = g[j, j:] * ri
        R[j, j] = d * ri
        g[j + 1:, j + 1:] -= np.outer(R[j, j + 1:], R[j, j + 1:])
    return R, ok


def triu_inverse(R, base=8):
    """Explicit inverse of an upper-triangular matrix by recursive doubling (diagonal base blocks by substitution, then
    X12 = -X11 (R12 X22) level by level), the order of operations of the kernel's trinv."""
    n = R.shape[0]
    X = np.zeros((n, n))
    for k in range(0, n, base):
        e = min(k + base, n)
        Rb = R[k:e, k:e]
        inv = np.zeros((e - k, e - k))
        for c in range(e - k):
            for i in range(c, -1, -1):
                s = 1.0 if i == c else 0.0
                s -= Rb[i, i + 1:c + 1] @ inv[i + 1:c + 1, c]
                inv[i, c] = s / Rb[i, i]
        X[k:e, k:e] = inv
    bs = base
    while bs < n:
        for o in range(0, n, 2 * bs):
            a, b, c = o, min(o + bs, n), min(o + 2 * bs, n)
            if b >= c:
                continue
            X[a:b, b:c] = -X[a:b, a:b] @ (R[a:b, b:c] @ X[b:c, b:c])
        bs *= 2
    return X


def kappa_estimate(R, X):
    """||D X||_F with D = diag(column norms of R) = diag(||p_j||): invariant under column scaling of the panel."""
    d = np.sqrt((R * R).sum(0))
    return float(np.sqrt(((d[:, None] * X) ** 2).sum()))


def signed_lu(W):
    """LU of the top block of E - Q S with S_j = -sign(pivot) chosen on the fly; returns (Wt, Sg, Ud):
    strict upper part of Wt = frozen rows U, strict lower part = W_ij^(j)."""
    Wt = np.array(W, dtype=np.float64, copy=True)
    n = Wt.shape[0]
    Sg, Ud = np.zeros(n), np.zeros(n)
    for j in range(n):
        w = Wt[j, j]
        Sg[j] = -1.0 if w > 0.0 else 1.0
        Ud[j] = 1.0 + abs(w)
        f = Sg[j] / Ud[j]
        Wt[j + 1:, j + 1:] += np.outer(f * Wt[j + 1:, j], Wt[j, j + 1:])
    return Wt, Sg, Ud


def wide_panel(P, kappa_max=KAPPA_MAX):
    """Returns (H, alpha, ok): H in the reference's storage; ok False when a guard refuses the panel (non-positive or
    non-finite Cholesky pivot, ||D R1^{-1}||_F > kappa_max, or the first pass left ||Q1'Q1 - I|| > 1/4) — the driver then
    redoes the panel with the 32-column chain.  The guards are invariant under column scaling."""
    P = np.array(P, dtype=np.float64)
    m, n = P.shape
    assert m >= n
    R1, ok = cholesky_upper(P.T @ P)
    if not ok:
        return None, None, False
    X1 = triu_inverse(R1)
    if not (kappa_estimate(R1, X1) <= kappa_max):
        return None, None, False
    Q1 = P @ X1
    G2 = Q1.T @ Q1
    if not np.all(np.abs(G2 - np.eye(n)) <= ORTH_MAX / n):
        return None, None, False
    R2, ok = cholesky_upper(G2)
    if not ok:
        return None, None, False
    X2 = triu_inverse(R2)
    Rt = np.triu(R2 @ R1)                                   # k_trimm128
    Wt, Sg, Ud = signed_lu(Q1[:n] @ X2)                     # k_vpk_rmul on the top chunks, k_hr128
    rsq = 1.0 / np.sqrt(Ud)
    sq = Ud * rsq
    cl = -Sg / Ud
    Rr = np.diag(sq) + (cl * sq)[:, None] * np.triu(Wt, 1)
    X3 = np.triu(X2 @ triu_inverse(Rr))                     # k_hr128 (inverse), k_trimm128 (product)
    H = np.zeros((m, n))
    H[n:] = Q1[n:] @ X3                                     # k_vpk_rmul with the output to user storage
    H[:n] = np.tril(Wt, -1) * rsq[None, :] + np.diag(-Sg * sq) + Sg[:, None] * np.triu(Rt, 1)
    alpha = Sg * np.diag(Rt)
    return H, alpha, True


def gram_T(H):
    """T' from the Gram matrix, as k_tinv does: T^{-1} = I + striu(V'V)."""
    n = H.shape[1]
    V = np.tril(H)
    return np.linalg.inv(np.eye(n) + np.triu(V.T @ V, 1)).T


def blocked_qr(A, nb=NB):
    """Right-looking blocked QR with wide panels (what qr_blocked does with option wide_panel=1 on aligned full panels)."""
    A = np.array(A, dtype=np.float64, copy=True)
    m, n = A.shape
    alpha = np.zeros(n)
    for c in range(0, n, nb):
        kb = min(nb, n - c)
        H, a, ok = wide_panel(A[c:, c:c + kb])
        if not ok:
            return None, None, c
        A[c:, c:c + kb] = H
        alpha[c:c + kb] = a
        if c + kb < n:
            V = np.tril(H)
            A[c:, c + kb:] -= V @ (gram_T(H) @ (V.T @ A[c:, c + kb:]))
    return A, alpha, -1
