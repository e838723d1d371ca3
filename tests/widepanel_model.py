"""numpy restatement of the 128-column panel chain (k_chol128 / k_gram2_finish / k_vpk_rmul / k_hr128 / k_trimm128 / k_trimm_z in
distributedhouseholderqr.jl_b200/csrc/dhqr_wide.cuh), stage by stage:

    G1 = P'P          R1 = chol(G1)     Z1 = blocked inverse operand of R1                Q1 = P R1^{-1} (through Z1)
    G2 = Q1'Q1        E = G2 - I        R2 = I + U, Z2 = I - U with U = striu(E) + diag(E)/2   (first order in E; a panel
                                        with max|E| > 1e-9 is refused: it did not pass the first pass well conditioned)
    Wt = Q1[:nb] R2^{-1}   signed LU of Wt (Householder reconstruction: Ballard, Demmel, Grigori, Jacquelin, Nguyen, Solomonik 2014)
    Rr = diag(sqrt(Ud)) (I + diag(cl) striu(U))      V[nb:] = Q1[nb:] (Rr R2)^{-1}      Rt = R2 R1

"Blocked inverse operand" Z of an upper triangular R, 32-column blocks: Z_bb = inv(R_bb), Z_ab = -R_ab inv(R_bb) for a < b, so
that X = P R^{-1} is the GEMM-shaped, row-local recurrence X_b = P_b Z_bb + sum_{a<b} X_a Z_ab.

The output is the reference's storage (S:127-135: v scaled to |v|^2 = 2 in the lower trapezoid including the diagonal, R above,
diag(R) in alpha) for a whole outer panel with three grid-wide reductions instead of one per column.

Test infrastructure only (tests/test_widepanel_model.py): pins on the CPU that this algorithm yields the reflectors of the
reference's column recurrences and that the guards refuse ill-conditioned panels.
"""
import numpy as np

NB = 128
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


BS = 32                  # block size of the inverse operand
FIRST_ORDER_MAX = 1.0e-9 # second pass: max|Q1'Q1 - I| up to which chol(I + E) is taken to first order in E


def inverse_operand(R):
    """Z_bb = inv(R_bb), Z_ab = -R_ab inv(R_bb) (a < b)."""
    n = R.shape[0]
    Z = np.zeros((n, n))
    for b in range(0, n, BS):
        D = triu_inverse(R[b:b + BS, b:b + BS], base=BS)
        Z[b:b + BS, b:b + BS] = D
        for a in range(0, b, BS):
            Z[a:a + BS, b:b + BS] = -(R[a:a + BS, b:b + BS] @ D)
    return Z


def solve_right(P, Z):
    """X = P R^{-1} through the inverse operand Z of R (k_vpk_rmul): X_b = P_b Z_bb + sum_{a<b} X_a Z_ab."""
    X = np.array(P, dtype=np.float64, copy=True)
    n = Z.shape[0]
    for b in range(0, n, BS):
        acc = P[:, b:b + BS] @ Z[b:b + BS, b:b + BS]
        for a in range(0, b, BS):
            acc = acc + X[:, a:a + BS] @ Z[a:a + BS, b:b + BS]
        X[:, b:b + BS] = acc
    return X


def kappa_estimate(R, Z):
    """sqrt(sum_b ||D_b Z_bb||_F^2 + sum_{a<b} ||Z_ab||_F^2), D = diag(column norms of R) = diag(||p_j||): invariant under
    column scaling of the panel; ~ the condition number of the column-normalised panel."""
    d = np.sqrt((R * R).sum(0))
    n = R.shape[0]
    s = 0.0
    for b in range(0, n, BS):
        s += ((d[b:b + BS, None] * Z[b:b + BS, b:b + BS]) ** 2).sum()
        for a in range(0, b, BS):
            s += (Z[a:a + BS, b:b + BS] ** 2).sum()
    return float(np.sqrt(s))


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


def second_pass(G2):
    """(R2, Z2, ok) from G2 = Q1'Q1 (k_gram2_finish): chol(I + E) and its inverse to first order in E = G2 - I; refuses the
    panel when max|E| > FIRST_ORDER_MAX (the neglected terms are O(n E^2)) — E is O(eps kappa^2), so a panel that passed the
    conditioning guard of the first pass is orders of magnitude below the bound."""
    n = G2.shape[0]
    E = G2 - np.eye(n)
    if not np.all(np.abs(E) <= FIRST_ORDER_MAX):
        return None, None, False
    U = np.triu(E, 1) + np.diag(np.diag(E)) / 2
    return np.eye(n) + U, np.eye(n) - U, True


def wide_panel(P, kappa_max=KAPPA_MAX):
    """Returns (H, alpha, ok): H in the reference's storage; ok False when a guard refuses the panel (non-positive or
    non-finite Cholesky pivot, conditioning estimate > kappa_max, or the first pass left max|Q1'Q1 - I| > 1e-9) — the driver then
    redoes the panel with the 32-column chain.  The guards are invariant under column scaling."""
    P = np.array(P, dtype=np.float64)
    m, n = P.shape
    assert m >= n and n % BS == 0
    R1, ok = cholesky_upper(P.T @ P)
    if not ok:
        return None, None, False
    Z1 = inverse_operand(R1)
    if not (kappa_estimate(R1, Z1) <= kappa_max):
        return None, None, False
    Q1 = solve_right(P, Z1)
    R2, Z2, ok = second_pass(Q1.T @ Q1)
    if not ok:
        return None, None, False
    Rt = np.triu(R2 @ R1)                                   # k_trimm128
    Wt, Sg, Ud = signed_lu(solve_right(Q1[:n], Z2))         # k_vpk_rmul on the top chunks, k_hr128
    rsq = 1.0 / np.sqrt(Ud)
    sq = Ud * rsq
    cl = -Sg / Ud
    Rr = np.diag(sq) + (cl * sq)[:, None] * np.triu(Wt, 1)
    Z23 = inverse_operand(np.triu(Rr @ R2))                 # k_trimm_z
    H = np.zeros((m, n))
    H[n:] = solve_right(Q1[n:], Z23)                        # k_vpk_rmul with the output to user storage
    H[:n] = np.tril(Wt, -1) * rsq[None, :] + np.diag(-Sg * sq) + Sg[:, None] * np.triu(Rt, 1)
    alpha = Sg * np.diag(Rt)
    return H, alpha, True


def reconstruction_T(H, Wt, Sg, Ud):
    """T' from the reconstruction (k_hr128 + k_trecon): V1 T V1' = E - Q S = L_lu U_lu  =>  T' = V1^{-1} (U_lu' D'^{-1}), with
    U_lu(j, j) = Ud_j, U_lu(j, k) = -Up(j, k) S_k (k > j) and D' = diag(v_jj)."""
    n = H.shape[1]
    sq = np.sqrt(Ud)
    MT = np.diag(-Sg * sq) + np.tril((np.triu(Wt, 1) * Sg[None, :] * (Sg / sq)[:, None]).T, -1)
    return np.linalg.solve(np.tril(H[:n]), MT)


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
