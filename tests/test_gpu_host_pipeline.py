"""dhqr_qr_host_f64 as a pipeline (S:311-315 with the matrix in host memory): the matrix goes up in column chunks, the
factorisation starts on the first one, later chunks join the trailing matrix through a catch-up with the reflectors of the
panels already finished.  Whatever the plan (chunk width, assumed link speed), every column must receive every reflector once
and in order: parity with the oracle, and the restart after a refused 128-column panel must still see a consistent matrix."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_H, TOL_A, TOL_RES = 1e-10, 1e-12, 1e-13


@pytest.fixture(scope="module")
def D():
    import dhqr_b200
    assert torch.cuda.is_available()
    return dhqr_b200


def host_qr(D, h, A0, nb=0):
    """pinned, column-major host copy of A0 -> dhqr_qr_host_f64 (truly asynchronous copies) -> (H, alpha) as numpy"""
    m, n = A0.shape
    hostA = torch.empty((n, m), dtype=torch.float64).pin_memory().t()
    hostA.copy_(torch.from_numpy(np.ascontiguousarray(A0)))
    alpha = torch.empty(n, dtype=torch.float64).pin_memory()
    D._lib.call("dhqr_qr_host_f64", h.raw, m, n, C.c_void_p(hostA.data_ptr()), m, C.c_void_p(alpha.data_ptr()), nb)
    return np.asfortranarray(hostA.numpy().copy()), alpha.numpy().copy()


PLANS = [  # (host_chunk, assumed link GB/s): slow link -> every join forced at the last moment (longest catch-up);
    (128, 1), (128, 100000), (256, 50), (512, 50), (384, 3), (0, 50)]   # fast link -> joins at step 0; 0 = one upload


@pytest.mark.parametrize("mn", [(4096, 2176), (3000, 1408), (2500, 1100)])
def test_every_plan_gives_the_reference_factorisation(D, oracle, coracle, mn):
    m, n = mn
    h = D.default_handle(0)
    A0 = coracle.fill_uniform(5, m, n)
    Href = A0.copy(order="F")
    Href, aref = coracle.qr(Href)
    try:
        for chunk, gbs in PLANS:
            h.set_option("host_chunk", chunk)
            h.set_option("host_h2d_gbs", gbs)
            H, a = host_qr(D, h, A0)
            assert np.abs(H - Href).max() < TOL_H, (chunk, gbs)
            assert np.abs(a - aref).max() < TOL_A * np.abs(aref).max(), (chunk, gbs)
            assert oracle.qr_residual(A0, H, a) < TOL_RES, (chunk, gbs)
    finally:
        h.set_option("host_chunk", 512)
        h.set_option("host_h2d_gbs", 50)


def test_narrower_panels_and_unblocked_through_the_host_entry(D, oracle, coracle):
    m, n = 2304, 1152
    h = D.default_handle(0)
    A0 = coracle.fill_uniform(6, m, n)
    Href = A0.copy(order="F")
    Href, aref = coracle.qr(Href)
    try:
        h.set_option("host_chunk", 128)
        for nb in (32, 64, 96, 1):
            H, a = host_qr(D, h, A0, nb)
            assert np.abs(H - Href).max() < TOL_H, nb
            assert np.abs(a - aref).max() < TOL_A * np.abs(aref).max(), nb
    finally:
        h.set_option("host_chunk", 512)


@pytest.mark.parametrize("gbs", [1, 100000])
def test_restart_inside_the_pipeline(D, oracle, coracle, gbs):
    # panel 5 (columns 640..767) nearly rank deficient: the wide chain refuses it while chunks are still joining; the catch-ups
    # behind it are gated like every other update, and the restart must find V_0..V_4 applied to every column right of it
    m, n = 3000, 1408
    h = D.default_handle(0)
    A0 = coracle.fill_uniform(14, m, n)
    A0[:, 700] = A0[:, 650] + 1e-11 * coracle.fill_uniform(15, m, 1)[:, 0]
    res = {}
    try:
        h.set_option("host_chunk", 128)
        h.set_option("host_h2d_gbs", gbs)
        for wide in (1, 0):
            h.set_option("wide_panel", wide)
            r0 = h.get_option("wide_redone")
            res[wide] = host_qr(D, h, A0)
            if wide:
                assert h.get_option("wide_redone") == r0 + 1
    finally:
        h.set_option("wide_panel", 1)
        h.set_option("host_chunk", 512)
        h.set_option("host_h2d_gbs", 50)
    for wide in (1, 0):
        assert oracle.qr_residual(A0, res[wide][0], res[wide][1]) < TOL_RES
    assert np.abs(res[1][1] - res[0][1]).max() < 1e-5 * np.abs(res[0][1]).max()
    Hr = A0.copy(order="F")
    Hr, ar = coracle.qr(Hr)
    assert np.abs(res[1][0][:, :640] - Hr[:, :640]).max() < TOL_H      # everything left of the refused panel: exact parity


# ---- Q'b / Qb with one right-hand side: GEMV sweep behind a batched T' (k_qt_dot / k_qt_axpy) ------------------------------
@pytest.mark.parametrize("mn", [(1024, 128), (1001, 37), (4400, 4000), (3000, 650), (8192, 1024)])
def test_vector_qt_sweep_against_the_block_update_and_the_oracle(D, oracle, coracle, mn):
    m, n = mn
    dev = torch.device("cuda:0")
    h = D.default_handle(0)
    A0 = coracle.fill_uniform(7, m, n)
    Href = A0.copy(order="F")
    Href, aref = coracle.qr(Href)
    b = oracle.np_uniform(8, m, 1)[:, 0].copy()
    qtb_ref = coracle.apply_qt(Href, b.copy()) if hasattr(coracle, "apply_qt") else None
    A = D.to_colmajor(A0, dev)
    H = D.qr_(A)
    res = {}
    try:
        for vec in (1, 0):
            h.set_option("qt_vec", vec)
            w = torch.from_numpy(b).to(dev)
            D.apply_qt_(w, A, h)
            res[vec] = w.cpu().numpy()
            back = w.clone()
            D.apply_q_(back, A, h)                                     # Q (Q'b) = b: the reverse sweep with T instead of T'
            assert np.abs(back.cpu().numpy() - b).max() < 1e-12 * np.abs(b).max() * np.sqrt(m), vec
    finally:
        h.set_option("qt_vec", 1)
    nb = np.linalg.norm(b)
    assert np.linalg.norm(res[1] - res[0]) < 1e-13 * nb
    if qtb_ref is not None:
        assert np.linalg.norm(res[1] - qtb_ref) < 1e-12 * nb           # ||Q'b - reference|| / ||b||  (BASELINE's second parity figure)
    x = D.ldiv(H, torch.from_numpy(b).to(dev)).cpu().numpy()
    xr = coracle.ldiv(Href, aref, b)
    assert np.abs(x - xr).max() < 1e-8 * np.abs(xr).max()
