"""CPU pin of the 128-column panel chain's algorithm (tests/widepanel_model.py restates the kernels stage by stage):
same reflectors as the reference's column recurrences (S:127-135, S:198-213), residual at rounding level for every panel the
guards accept, ill-conditioned / rank-deficient / non-finite panels refused, guards invariant under column scaling."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import widepanel_model as W   # noqa: E402


def colres(oracle, P, H, a):
    R = oracle.reconstruct(np.asfortranarray(H), a) - P
    return float((np.linalg.norm(R, axis=0) / np.linalg.norm(P, axis=0)).max())


@pytest.mark.parametrize("m", [128, 130, 256, 1024, 4096])
def test_same_reflectors_as_the_reference_recurrences(oracle, m):
    P = oracle.np_uniform(3, m, 128)
    H, a, ok = W.wide_panel(P)
    assert ok
    Hr, ar = oracle.np_qr(P)
    assert np.abs(H - Hr).max() < 1e-12
    assert np.abs(a - ar).max() < 1e-13 * np.abs(ar).max()
    assert oracle.qr_residual(P, np.asfortranarray(H), a) < 3e-15
    V = np.tril(H)
    assert np.abs((V * V).sum(0) - 2.0).max() < 1e-13            # |v|^2 = 2 (S:131-135)


def test_compact_wy_factor_from_the_reconstruction(oracle):
    P = oracle.np_uniform(8, 2048, 128)
    R1, _ = W.cholesky_upper(P.T @ P)
    Q1 = W.solve_right(P, W.inverse_operand(R1))
    R2, Z2, ok = W.second_pass(Q1.T @ Q1)
    Wt, Sg, Ud = W.signed_lu(W.solve_right(Q1[:128], Z2))
    H, a, ok = W.wide_panel(P)
    Tt = W.reconstruction_T(H, Wt, Sg, Ud)
    assert np.abs(Tt - W.gram_T(H)).max() < 1e-14 and np.abs(np.triu(Tt, 1)).max() < 1e-15


def test_blocked_sweep_matches_oracle(oracle):
    A = oracle.np_uniform(0, 1100, 1024)
    H, a, bad = W.blocked_qr(A)
    assert bad < 0
    Hr, ar = oracle.np_qr(A)
    assert np.abs(H - Hr).max() < 1e-11
    assert np.abs(a - ar).max() < 1e-12 * np.abs(ar).max()
    assert oracle.qr_residual(A, np.asfortranarray(H), a) < 5e-15


def ill_conditioned_panels(oracle):
    rng = np.random.default_rng(5)
    m, n = 1024, 128
    U, _ = np.linalg.qr(rng.standard_normal((m, n)))
    Vt, _ = np.linalg.qr(rng.standard_normal((n, n)))
    P0 = oracle.np_uniform(8, m, n)
    for kappa in (1e1, 1e2, 1e3, 1e4, 1e6, 1e8, 1e12):
        yield f"geo{kappa:.0e}", (U * np.logspace(0, -np.log10(kappa), n)) @ Vt.T
        yield f"one{kappa:.0e}", (U * np.r_[np.ones(n - 1), 1.0 / kappa]) @ Vt.T
    for eps in (1e-1, 1e-2, 1e-3, 1e-4, 1e-6, 1e-9):       # two nearly identical columns on top of the rank-one mean of U[0,1)
        for c in (9, 99):
            Q = P0.copy()
            Q[:, c] = Q[:, 2] + eps * oracle.np_uniform(10, m, 1)[:, 0]
            yield f"dup{c}_{eps:.0e}", Q


def test_guards_keep_every_accepted_panel_backward_stable(oracle):
    accepted = 0
    for name, P in ill_conditioned_panels(oracle):
        for scaled in (False, True):
            Q = P * np.logspace(-6, 6, 128)[None, :] if scaled else P
            H, a, ok = W.wide_panel(Q)
            H1, a1, ok1 = W.wide_panel(P)
            assert ok == ok1, name                                     # the guards are invariant under column scaling
            if ok:
                accepted += 1
                assert colres(oracle, Q, H, a) < 5e-14, name           # column by column, not just in the Frobenius norm
    assert accepted >= 12                                              # ... and they do not refuse everything
    for name, P in ill_conditioned_panels(oracle):                     # well-conditioned panels must pass
        if name in ("geo1e+01", "one1e+01", "geo1e+02", "dup9_1e-01", "dup99_1e-01", "dup9_1e-02"):
            assert W.wide_panel(P)[2], name
        if name.endswith("1e+12") or name.endswith("1e-09"):
            assert not W.wide_panel(P)[2], name


def test_the_two_guards_overlap(oracle):
    # explicit 32 x 32 inverses lose accuracy in proportion to the conditioning estimate; the same panels lose orthogonality in
    # the first pass (E ~ eps kappa^2), so even with the conditioning guard opened up the second pass refuses them
    P = oracle.np_uniform(8, 1024, 128)
    Q = P.copy()
    Q[:, 9] = Q[:, 2] + 1e-5 * oracle.np_uniform(10, 1024, 1)[:, 0]
    assert not W.wide_panel(Q)[2] and not W.wide_panel(Q, kappa_max=1e30)[2]
    Q = P.copy()
    Q[:, 9] = Q[:, 2] + 2e-3 * oracle.np_uniform(10, 1024, 1)[:, 0]                # est ~ 1.5e3: refused by default ...
    assert not W.wide_panel(Q)[2]
    H, a, ok = W.wide_panel(Q, kappa_max=1e30)                                     # ... and still accurate if let through
    assert ok and colres(oracle, Q, H, a) < 5e-14


def test_degenerate_panels_are_refused():
    P = np.random.default_rng(0).random((512, 128))
    for bad in ("zero", "dup", "nan", "inf"):
        Q = P.copy()
        if bad == "zero":
            Q[:, 17] = 0.0
        elif bad == "dup":
            Q[:, 40] = Q[:, 3]
        elif bad == "nan":
            Q[100, 5] = np.nan
        else:
            Q[7, 99] = np.inf
        with np.errstate(all="ignore"):
            assert not W.wide_panel(Q)[2]


def test_inverse_operand_solves_from_the_right():
    rng = np.random.default_rng(1)
    R = np.triu(rng.standard_normal((128, 128))) + 12 * np.eye(128)
    X = W.triu_inverse(R)
    assert np.abs(X @ R - np.eye(128)).max() < 1e-14 and np.abs(np.tril(X, -1)).max() == 0.0
    P = rng.standard_normal((300, 128))
    Y = W.solve_right(P, W.inverse_operand(R))
    assert np.abs(Y @ R - P).max() < 1e-13 * np.abs(P).max() * 128


def test_second_pass_first_order_factor():
    # chol(I + E) = I + U + O(E^2), inverse I - U + O(E^2) with U = striu(E) + diag(E)/2: exact to rounding below 1e-9
    rng = np.random.default_rng(2)
    for mag in (1e-15, 1e-12, 1e-9 / 2):
        E = rng.uniform(-mag, mag, (128, 128))
        E = (E + E.T) / 2
        R2, Z2, ok = W.second_pass(np.eye(128) + E)
        assert ok
        Rc, _ = W.cholesky_upper(np.eye(128) + E)
        assert np.abs(R2 - Rc).max() < 128 * mag * mag + 4e-16
        assert np.abs(Z2 @ R2 - np.eye(128)).max() < 128 * mag * mag + 4e-16
    assert not W.second_pass(np.eye(128) + 1e-6 * np.ones((128, 128)))[2]      # beyond first order: refused, not approximated
    assert not W.second_pass(np.full((128, 128), np.nan))[2]


def test_panels_inside_the_conditioning_guard_stay_far_below_the_first_order_bound(oracle):
    # E = Q1'Q1 - I is O(eps kappa^2): with est <= 1e3 it never comes near 1e-9 on tall panels
    worst = 0.0
    for m in (128, 200, 1024, 8192):
        P = oracle.np_uniform(4, m, 128)
        R1, ok = W.cholesky_upper(P.T @ P)
        Q1 = W.solve_right(P, W.inverse_operand(R1))
        worst = max(worst, np.abs(Q1.T @ Q1 - np.eye(128)).max())
    assert worst < 1e-10
