"""GPU parity tests of the 128-column panel chain (dhqr_wide.cuh): CholeskyQR2 + Householder reconstruction of a whole
outer panel, its on-device guards, and the restart of a factorisation after a refused panel.  Same tolerances as
tests/test_gpu_parity.py; the numpy restatement of the stages is tests/widepanel_model.py."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import widepanel_model as W   # noqa: E402

pytestmark = pytest.mark.gpu
TOL_H, TOL_A, TOL_RES = 1e-10, 1e-12, 1e-13


@pytest.fixture(scope="module")
def D():
    import dhqr_b200
    assert torch.cuda.is_available()
    return dhqr_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def vp(t):
    return C.c_void_p(t.data_ptr())


def sp():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def wide_panel(D, dev, P, lda=None):
    h = D.default_handle(0)
    rows = P.shape[0]
    dP = D.colmajor_empty(rows, 128, dev, lda=lda or rows)
    dP.copy_(torch.from_numpy(P))
    dal = torch.zeros(128, dtype=torch.float64, device=dev)
    refused = C.c_int(-1)
    D._lib.call("dhqr_k_wide_panel_f64", h.raw, rows, vp(dP), lda or rows, vp(dal), C.byref(refused), sp())
    torch.cuda.synchronize()
    return dP.cpu().numpy(), dal.cpu().numpy(), refused.value


XL_OFF = [0, 1152, 3328, 6528]
XL_ELEMS = 10752


def unpack_operand(zl):
    """The rmul operand layout -> dense upper block-triangular 128 x 128 (rows k < 32 (b + 1) of block column b)."""
    Z = np.zeros((128, 128))
    for b in range(4):
        ld = 32 * (b + 1) + 4
        blk = zl[XL_OFF[b]:XL_OFF[b] + 32 * ld].reshape(32, ld)      # [n % 32][k]
        Z[:32 * (b + 1), 32 * b:32 * b + 32] = blk[:, :32 * (b + 1)].T
    return Z


def stages(D, dev):
    """R1, R2, Rt, Rr, MT (plain 128 x 128, column-major), the inverse operands Z1, Z2, Z23 and T' left by the last wide panel."""
    h = D.default_handle(0)
    buf = torch.zeros(5 * 128 * 128 + 3 * XL_ELEMS, dtype=torch.float64, device=dev)
    D._lib.call("dhqr_debug_copy_f64", h.raw, b"wide", vp(buf), buf.numel(), sp())
    torch.cuda.synchronize()
    a = buf.cpu().numpy()
    out = {k: a[i * 16384:(i + 1) * 16384].reshape(128, 128).T.copy() for i, k in enumerate(["R1", "R2", "Rt", "Rr", "MT"])}
    for i, k in enumerate(["Z1", "Z2", "Z23"]):
        out[k] = unpack_operand(a[5 * 16384 + i * XL_ELEMS:5 * 16384 + (i + 1) * XL_ELEMS])
    lin = torch.zeros(128 * 128, dtype=torch.float64, device=dev)
    D._lib.call("dhqr_debug_copy_f64", h.raw, b"linv", vp(lin), lin.numel(), sp())
    torch.cuda.synchronize()
    out["Tt"] = lin.cpu().numpy().reshape(128, 128).T.copy()
    return out


def test_stage_outputs_against_the_numpy_model(D, dev, oracle):
    # pins each kernel of the chain separately: a failure here names the stage
    rows = 1024
    P = oracle.np_uniform(31, rows, 128)
    H, al, refused = wide_panel(D, dev, P)
    assert refused == 0
    st = stages(D, dev)
    R1, ok = W.cholesky_upper(P.T @ P)
    Z1 = W.inverse_operand(R1)
    Q1 = W.solve_right(P, Z1)
    R2, Z2, ok2 = W.second_pass(Q1.T @ Q1)
    assert ok and ok2
    sc = np.abs(R1).max()
    assert np.abs(st["R1"] - R1).max() < 1e-11 * sc, "k_gemm_vta Gram / k_chol128"
    assert np.abs(np.tril(st["R1"], -1)).max() == 0.0
    assert np.abs(st["Z1"] - W.inverse_operand(st["R1"])).max() < 1e-10 * np.abs(Z1).max(), "k_chol128: inverse operand"
    assert np.abs(st["R2"] - R2).max() < 1e-11, "k_vpk_rmul (Z1) / Gram / k_gram2_finish"
    assert np.abs(st["Z2"] - Z2).max() < 1e-11, "k_gram2_finish: inverse operand"
    assert np.abs(st["Rt"] - np.triu(st["R2"] @ st["R1"])).max() < 1e-12 * sc, "k_trimm128"
    Wt, Sg, Ud = W.signed_lu(W.solve_right(Q1[:128], Z2))
    rsq = 1.0 / np.sqrt(Ud)
    Rr = np.diag(Ud * rsq) + ((-Sg / Ud) * Ud * rsq)[:, None] * np.triu(Wt, 1)
    assert np.abs(st["Rr"] - Rr).max() < 1e-10, "k_hr128 (signed LU, Rr)"
    assert np.abs(st["Z23"] - W.inverse_operand(np.triu(st["Rr"] @ st["R2"]))).max() < 1e-12, "k_trimm_z"
    Hm, am, okm = W.wide_panel(P)
    assert okm and np.abs(H - Hm).max() < 1e-9 and np.abs(al - am).max() < 1e-9 * np.abs(am).max()
    assert oracle.qr_residual(P, np.asfortranarray(H), al) < TOL_RES
    # k_trecon: T' from the reconstruction == (I + stril(V'V))^{-1} (what k_tinv computes from the Gram matrix)
    assert np.abs(st["Tt"] - W.gram_T(H)).max() < 1e-13, "k_hr128 (MT) / k_trecon"
    assert np.abs(np.triu(st["Tt"], 1)).max() == 0.0


def test_loss_of_orthogonality_beyond_first_order_refuses_the_panel(D, dev, oracle):
    # with the conditioning guard opened up, a nearly dependent pair of columns leaves max|Q1'Q1 - I| ~ 1e-7 after the first
    # pass: the second pass (first order in that quantity) must refuse, exactly like the model
    h = D.default_handle(0)
    rows = 1024
    P = oracle.np_uniform(31, rows, 128)
    P[:, 70] = P[:, 5] + 1e-4 * oracle.np_uniform(32, rows, 1)[:, 0]
    h.set_option("wide_kappa", 10 ** 6)
    try:
        H, al, refused = wide_panel(D, dev, P)
    finally:
        h.set_option("wide_kappa", 1000)
    assert refused == 1 and not W.wide_panel(P, kappa_max=1e6)[2]
    assert np.array_equal(H, P)


@pytest.mark.parametrize("rows", [128, 129, 130, 192, 200, 1000, 1024, 4097, 32768, 65536])
def test_wide_panel_kernel(D, dev, oracle, coracle, rows):
    P = coracle.fill_uniform(7, rows, 128)
    Href = P.copy(order="F")
    Href, aref = coracle.qr(Href)
    lda = rows + (3 if rows % 2 else 0)                        # odd leading dimension on the ragged sizes
    H, al, refused = wide_panel(D, dev, P, lda)
    assert refused == 0
    assert np.abs(H - Href).max() < TOL_H
    assert np.abs(al - aref).max() < TOL_A * np.abs(aref).max()
    assert oracle.qr_residual(P, np.asfortranarray(H), al) < TOL_RES
    # the packed V block the trailing GEMMs read == tril(H), zero padded
    h = D.default_handle(0)
    vrows = (rows + 127) // 128 * 128
    buf = torch.zeros(vrows // 64 * 128 * 68, dtype=torch.float64, device=dev)
    D._lib.call("dhqr_debug_copy_f64", h.raw, b"vpk", vp(buf), buf.numel(), sp())
    torch.cuda.synchronize()
    V = buf.cpu().numpy().reshape(vrows // 64, 128, 68)[:, :, :64].transpose(0, 2, 1).reshape(vrows, 128)
    assert np.array_equal(V[:rows], np.tril(H)) and not V[rows:].any()


def test_guards_refuse_on_the_device_and_leave_the_panel_alone(D, dev, oracle):
    rows = 2048
    P = oracle.np_uniform(8, rows, 128)
    cases = {}
    Pi = P.copy()
    Pi[:, 77] = Pi[:, 3] + 1e-10 * oracle.np_uniform(9, rows, 1)[:, 0]       # kappa ~ 1e10: Cholesky or the Q1'Q1 guard
    cases["dependent"] = Pi
    Pz = P.copy()
    Pz[:, 100] = 0.0
    cases["zero column"] = Pz
    Pn = P.copy()
    Pn[5, 5] = np.nan
    cases["nan"] = Pn
    for name, Q in cases.items():
        H, al, refused = wide_panel(D, dev, Q)
        assert refused == 1, name
        assert np.array_equal(H, Q, equal_nan=True), name                     # nothing was written to the caller's panel
    # column scaling alone is no reason to refuse (the guards are scale invariant), and the result stays backward stable
    Ps = P * np.logspace(-6, 6, 128)[None, :]
    H, al, refused = wide_panel(D, dev, Ps)
    assert refused == 0
    R = oracle.reconstruct(np.asfortranarray(H), al) - Ps
    assert (np.linalg.norm(R, axis=0) / np.linalg.norm(Ps, axis=0)).max() < 1e-13
    # moderately ill-conditioned: accepted or refused (the guard on ||D R1^{-1}||_F decides), never inaccurate; the device
    # decision is the numpy model's
    for eps in (1e-1, 1e-2, 1e-3, 1e-5):
        Pm = P.copy()
        Pm[:, 9] = Pm[:, 2] + eps * oracle.np_uniform(10, rows, 1)[:, 0]
        H, al, refused = wide_panel(D, dev, Pm)
        assert refused == (0 if W.wide_panel(Pm)[2] else 1), eps
        if not refused:
            R = oracle.reconstruct(np.asfortranarray(H), al) - Pm
            assert (np.linalg.norm(R, axis=0) / np.linalg.norm(Pm, axis=0)).max() < 5e-14, eps
    assert wide_panel(D, dev, P * 1.0)[2] == 0


def test_restart_after_a_refused_panel(D, dev, oracle, coracle):
    # second outer panel nearly rank deficient: the wide chain refuses it on the device, everything behind it is skipped,
    # and qr! redoes the factorisation from that panel with the 32-column chain.  Same answer as with the wide chain off.
    h = D.default_handle(0)
    m, n = 3000, 640
    A0 = coracle.fill_uniform(12, m, n)
    A0[:, 200] = A0[:, 150] + 1e-11 * coracle.fill_uniform(13, m, 1)[:, 0]
    res = {}
    try:
        for wide in (1, 0):
            h.set_option("wide_panel", wide)
            r0, w0 = h.get_option("wide_redone"), h.get_option("wide_panels")
            A = D.to_colmajor(A0, dev)
            H = D.qr_(A)
            torch.cuda.synchronize()
            res[wide] = (A.cpu().numpy(), H.α.cpu().numpy())
            if wide:
                assert h.get_option("wide_redone") == r0 + 1
                assert h.get_option("wide_panels") > w0 + 3               # panels 0, 2, 3, 4 did go through the wide chain
    finally:
        h.set_option("wide_panel", 1)
    for wide in (1, 0):
        Hx, ax = res[wide]
        assert oracle.qr_residual(A0, np.asfortranarray(Hx), ax) < TOL_RES
    assert np.abs(res[1][1] - res[0][1]).max() < 1e-5 * np.abs(res[0][1]).max()    # alpha_200 ~ 1e-11: limited by kappa
    Hr = A0.copy(order="F")
    Hr, ar = coracle.qr(Hr)
    assert np.abs(res[1][0][:, :128] - Hr[:, :128]).max() < TOL_H                  # the well-conditioned leading panel: exact parity


@pytest.mark.parametrize("mn", [(880, 800), (2200, 2000), (8192, 1024)])
def test_narrow_chain_keeps_its_parity(D, dev, oracle, coracle, mn):
    # the 32-column chain is the fallback of the wide chain and the only path for ragged panels: keep it covered at sizes
    # where the default now picks the wide chain
    h = D.default_handle(0)
    m, n = mn
    A0 = coracle.fill_uniform(0, m, n)
    Href = A0.copy(order="F")
    Href, aref = coracle.qr(Href)
    try:
        h.set_option("wide_panel", 0)
        A = D.to_colmajor(A0, dev)
        H = D.qr_(A)
        torch.cuda.synchronize()
    finally:
        h.set_option("wide_panel", 1)
    assert np.abs(A.cpu().numpy() - Href).max() < TOL_H
    assert np.abs(H.α.cpu().numpy() - aref).max() < TOL_A * np.abs(aref).max()


def test_wide_chain_is_what_runs_at_baseline_shapes(D, dev):
    h = D.default_handle(0)
    w0, r0 = h.get_option("wide_panels"), h.get_option("wide_redone")
    A = D.colmajor_empty(8192, 1024, dev)
    D.fill_uniform_(A, 0)
    D.qr_(A)
    torch.cuda.synchronize()
    assert h.get_option("wide_panels") == w0 + 8 and h.get_option("wide_redone") == r0
