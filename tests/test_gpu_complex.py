"""GPU parity tests of the ComplexF64 path (the reference's second element type, test/runtests.jl:43; S:9, S:51-59, S:162-196)
against the complex oracle oracle/dhqr_oracle.py:np_qr_c / np_apply_qt_c / np_backsolve_c (pinned on zgeqrf and on the
reference's normal-equation property in tests/test_oracle.py).  Same tolerances as the Float64 tests."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL_H, TOL_A, TOL_QTB, TOL_RES = 1e-10, 1e-12, 1e-12, 1e-13


@pytest.fixture(scope="module")
def D():
    import dhqr_b200
    assert torch.cuda.is_available()
    return dhqr_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def crand(oracle, seed, m, n):
    """rand(ComplexF64, m, n) (T:45): both parts U[0,1) from the counter-based generator."""
    return oracle.np_uniform(seed, m, n) + 1j * oracle.np_uniform(seed + 1000, m, n)


def test_partialdot_complex_suffixes(D, dev):
    # test/partialdot.jl:11-22: N = 1..20, random ComplexF64, every suffix: partialdot(a, b, i:N) ~ dot(a[i:end], b[i:end])
    g = torch.Generator().manual_seed(0)
    for N in range(1, 21):
        a = torch.complex(torch.rand(N, dtype=torch.float64, generator=g), torch.rand(N, dtype=torch.float64, generator=g)).to(dev)
        b = torch.complex(torch.rand(N, dtype=torch.float64, generator=g), torch.rand(N, dtype=torch.float64, generator=g)).to(dev)
        for i in range(N):
            ref = complex(torch.vdot(a[i:], b[i:]).item())            # conjugating dot, like LinearAlgebra.dot
            got = D.partialdot(a, b, range(i, N))
            assert abs(got - ref) <= 1e-13 * max(1.0, abs(ref))


def test_alphafactor_complex(D):
    for x in (1 + 1j, -2 + 0.5j, 3j, -1.0 + 0j, 0j):
        assert abs(D.alphafactor(x) - (-np.exp(1j * np.angle(x)))) < 1e-15


# the reference's own sizes (T:42) as far as the numpy oracle finishes in seconds, plus ragged shapes and a leading dimension
@pytest.mark.parametrize("mn", [(110, 100), (220, 200), (440, 400), (880, 800), (1100, 1000), (257, 97), (70, 64), (65, 65),
                                (300, 1), (1, 1)])
def test_complex_qr_and_solve_against_oracle(D, dev, oracle, mn):
    m, n = mn
    A0 = crand(oracle, 0, m, n)
    b = crand(oracle, 7, m, 1)[:, 0].copy()
    Href, aref = oracle.np_qr_c(A0)
    A = D.colmajor_empty(m, n, dev, lda=m + (m % 2), dtype=torch.complex128)
    A.copy_(torch.from_numpy(A0))
    H = D.qr_(A)
    Hg, ag = A.cpu().numpy(), H.α.cpu().numpy()
    assert H.A is A and ag.dtype == np.complex128
    assert np.abs(Hg - Href).max() < TOL_H
    assert np.abs(ag - aref).max() < TOL_A * np.abs(aref).max()
    assert np.linalg.norm(oracle.reconstruct_c(Hg, ag) - A0) < TOL_RES * np.linalg.norm(A0)
    V = np.tril(Hg)
    assert np.abs((np.abs(V) ** 2).sum(0) - 2.0).max() < 1e-12                 # |v|^2 = 2 (S:131-135)
    bt = torch.from_numpy(b).to(dev)
    qtb = D.apply_qt_(bt.clone(), A).cpu().numpy()
    assert np.linalg.norm(qtb - oracle.np_apply_qt_c(Href, b)) < TOL_QTB * np.linalg.norm(b)
    x = D.ldiv(H, bt).cpu().numpy()
    assert torch.equal(bt.cpu(), torch.from_numpy(b))                           # \ leaves b alone (S:318)
    xr = oracle.np_ldiv_c(Href, aref, b)
    assert np.abs(x - xr).max() < 1e-9 * max(1.0, np.abs(xr).max())
    if n > 1:
        xs = np.linalg.lstsq(A0, b, rcond=None)[0]                              # T:49: the stdlib solution
        ne = lambda z: np.linalg.norm(A0.conj().T @ (A0 @ z) - A0.conj().T @ b)
        assert ne(x) < max(8 * ne(xs), 1.5 * ne(xr))                            # T:62 (see test_gpu_parity for the max)


def test_complex_multiple_right_hand_sides_and_zero_pivot(D, dev, oracle):
    m, n, k = 500, 130, 3
    A0 = crand(oracle, 3, m, n)
    A0[0, 0] = 0.0                                                              # angle(0) = 0: alphafactor = -1 (S:9)
    Href, aref = oracle.np_qr_c(A0)
    A = D.to_colmajor(A0, dev)
    H = D.qr_(A)
    assert np.abs(A.cpu().numpy() - Href).max() < TOL_H
    assert aref[0].real < 0 and abs(aref[0].imag) < 1e-15 and abs(H.α[0].item() - aref[0]) < 1e-12 * abs(aref[0])
    B0 = crand(oracle, 11, m, k)
    X = D.ldiv(H, torch.from_numpy(B0).to(dev)).cpu().numpy()
    for j in range(k):
        xr = oracle.np_ldiv_c(Href, aref, B0[:, j].copy())
        assert np.abs(X[:, j] - xr).max() < 1e-9 * np.abs(xr).max()


def test_complex_full_size_properties(D, dev):
    # T:42's largest size through size-independent properties (the numpy oracle would take minutes): Q R = A, |v|^2 = 2,
    # ||Q^H b|| = ||b||, normal equations vs the stdlib (cuSOLVER) solution
    m, n = 4400, 4000
    g = torch.Generator(device=dev).manual_seed(5)
    A0 = D.colmajor_empty(m, n, dev, dtype=torch.complex128)
    A0.copy_(torch.complex(torch.rand(m, n, dtype=torch.float64, device=dev, generator=g),
                           torch.rand(m, n, dtype=torch.float64, device=dev, generator=g)))
    A = A0.clone()
    H = D.qr_(A)
    R = torch.zeros(m, n, dtype=torch.complex128, device=dev)
    R[:n] = torch.triu(A[:n], 1) + torch.diag(H.α)
    for k in range(((n - 1) // 64) * 64, -1, -64):                              # Q R = H_1 (... H_n R), 64 reflectors at a time
        kb = min(64, n - k)
        V = torch.tril(A[k:, k:k + kb])
        Tinv = torch.eye(kb, dtype=torch.complex128, device=dev) + torch.triu(V.conj().T @ V, 1)
        R[k:] -= V @ torch.linalg.solve_triangular(Tinv, V.conj().T @ R[k:], upper=True)
    assert float(torch.linalg.norm(R - A0) / torch.linalg.norm(A0)) < TOL_RES
    assert float(((torch.tril(A).abs() ** 2).sum(0) - 2.0).abs().max()) < 1e-12
    b = torch.complex(torch.rand(m, dtype=torch.float64, device=dev, generator=g), torch.rand(m, dtype=torch.float64, device=dev, generator=g))
    qtb = D.apply_qt_(b.clone(), A)
    assert abs(float(torch.linalg.norm(qtb) / torch.linalg.norm(b)) - 1.0) < 1e-13
    x = D.ldiv(H, b)
    xs = torch.linalg.lstsq(A0, b.unsqueeze(1)).solution[:, 0]
    ne = lambda z: float(torch.linalg.norm(A0.conj().T @ (A0 @ z) - A0.conj().T @ b))
    assert ne(x) < 8 * ne(xs)


def test_complex_rejects_what_it_does_not_cover(D, dev):
    A = D.colmajor_empty(8, 4, dev, dtype=torch.complex128)
    with pytest.raises(TypeError):
        D.apply_q_(torch.zeros(8, dtype=torch.complex128, device=dev), A)
    with pytest.raises(TypeError):
        D.qr_(D.colmajor_empty(8, 4, dev).to(torch.float32))
    with pytest.raises(D._lib.DhqrError) as e:
        D.qr_(D.colmajor_empty(3, 5, dev, dtype=torch.complex128))              # n > m
    assert e.value.code == -3
