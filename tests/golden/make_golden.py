"""Generates tests/golden/*.npz.

The reference is Julia (not installed) and ships no golden vectors (its tests draw from Julia's RNG,
test/runtests.jl:6,45-46), so these fixtures pin the *restated* algorithm instead: inputs come from the
counter-based generator (oracle/dhqr_oracle.py:np_uniform), outputs from the pure-numpy twin of
S:122-148/S:198-213/S:232-242/S:256-282, and each case is cross-checked here against LAPACK dgeqrf mapped
into the reference's storage format before it is written.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import dhqr_oracle as O  # noqa: E402

CASES = [(12, 5, 0), (40, 33, 1), (110, 100, 0), (130, 64, 2), (257, 97, 3)]   # (m, n, seed); 110x100 = T:42's first size

for m, n, seed in CASES:
    A = O.np_uniform(seed, m, n)
    b = O.np_uniform(seed + 1000, m, 1)[:, 0].copy()
    H, alpha = O.np_qr(A)
    qtb = O.np_apply_qt(H, b)
    x = O.np_backsolve(H, alpha, qtb)
    Hl, al = O.lapack_qr_refformat(A)
    assert np.abs(H - Hl).max() < 1e-12 and np.abs(alpha - al).max() < 1e-12, (m, n)
    xl = O.lapack_lstsq(A, b)
    assert np.abs(x - xl).max() < 1e-9 * max(1.0, np.abs(xl).max()), (m, n)
    np.savez_compressed(os.path.join(HERE, f"qr_{m}x{n}_seed{seed}.npz"), A=A, b=b, H=H, alpha=alpha, qtb=qtb, x=x)
    print(f"wrote qr_{m}x{n}_seed{seed}.npz  resid={O.qr_residual(A, H, alpha):.2e}")
