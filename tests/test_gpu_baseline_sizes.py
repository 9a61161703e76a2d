"""GPU parity at the sizes BASELINE.json names (configs[1], configs[2]) against LAPACK itself, not the restatement:
   * N = 10000 random: eigen_trd (d, |e|) vs SciPy/OpenBLAS dsytrd('U'); eigen_s eigenvalues vs dsterf on LAPACK's
     tridiagonal matrix; residual / orthogonality of the eigenvectors with cuBLAS FP64 GEMMs (torch), an
     independent checker;
   * N = 20000 mode 'N' (eigenvalues only: scaling + eigen_trd + eigen_bisect) vs cuSOLVER dsyevd
     (torch.linalg.eigvalsh on the same device);
   * Frank matrices through eigen_prd: the band entries the first reflector fixes, against the oracle;
   * Helmert families 7, 9 and 10 (benchmark/W.dat) through both drivers, w_test gates.
Tolerance everywhere: 10 n eps |A|_F (BASELINE.json north_star); gates <= 10 for the eigenvectors."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

F = lambda x: np.array(x, order="F", copy=True)


@pytest.fixture(scope="module")
def big10k():
    from scipy.linalg import eigvalsh_tridiagonal, lapack
    n = 10000
    a = O.mat_set(n, O.MAT_RANDOM)                 # full symmetric matrix, benchmark/mat_set.f type 2
    nrm = np.linalg.norm(a)
    _, d, e, _, info = lapack.dsytrd(F(a), lower=0)
    assert info == 0
    return n, a, nrm, d, e, eigvalsh_tridiagonal(d, e)


def test_trd_n10000_matches_lapack_dsytrd(ee, big10k):
    n, a, nrm, d_l, e_l, _ = big10k
    ag = F(a)
    d, e = ee.eigen_trd(n, ag, 48)
    tol = 10 * n * O.EPS * nrm
    assert e[0] == 0.0
    assert np.abs(d - d_l).max() <= tol
    # same sign convention as dsytrd('U') except e(2) (src/eigen_trd_t8.F:190-199): compare |e|
    assert np.abs(np.abs(e[1:]) - np.abs(e_l)).max() <= tol
    assert np.abs(e[2:] - e_l[1:]).max() <= tol     # ... and the signs do agree from e(3) on


def test_eigen_s_n10000_all_pairs_vs_lapack(ee, big10k):
    import torch
    n, a, nrm, _, _, w_l = big10k
    w, z = np.zeros(n), np.zeros((n, n), order="F")
    ee.eigen_s(n, F(a), w, z, m_forward=48, m_backward=128, mode="A")
    tol = 10 * n * O.EPS * nrm
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - w_l).max() <= tol
    dev = torch.device("cuda:0")
    A = torch.from_numpy(a).to(dev)                               # symmetric: layout does not matter
    Z = torch.from_numpy(np.ascontiguousarray(z.T)).to(dev).T     # n x n, column-major memory
    W = torch.from_numpy(w).to(dev)
    res = torch.linalg.norm(A @ Z - Z * W[None, :]).item() / (n * O.EPS * nrm)
    orth = torch.linalg.norm(Z.T @ Z - torch.eye(n, dtype=torch.float64, device=dev)).item() / (n * O.EPS)
    assert res <= 10 and orth <= 10, (res, orth)


def test_eigen_sx_n10000_eigenvalues_vs_lapack(ee, big10k):
    n, a, nrm, _, _, w_l = big10k
    w, z = np.zeros(n), np.zeros((n, n), order="F")
    ee.eigen_sx(n, F(a), w, z, m_forward=48, m_backward=128, mode="A")
    assert np.abs(w - w_l).max() <= 10 * n * O.EPS * nrm
    res, orth = ee_ev_test(ee, n, a, w, z)
    assert res <= 10 and orth <= 10, (res, orth)


def ee_ev_test(ee, n, a, w, z):
    """benchmark/ev_test.f on the device through the library's own GEMM (cross-checked against cuBLAS above)."""
    import torch
    dev = torch.device("cuda:0")
    A = torch.from_numpy(a).to(dev)
    Z = torch.from_numpy(np.ascontiguousarray(z.T)).to(dev)
    W = torch.from_numpy(w).to(dev)
    ee.sync()
    return ee.ev_test_dev(n, n, A.data_ptr(), n, W.data_ptr(), Z.data_ptr(), n)


def test_eigen_s_n20000_mode_n_vs_cusolver(ee):
    """BASELINE configs[2]: N = 20000, eigenvalues only."""
    import torch
    n = 20000
    dev = torch.device("cuda:0")
    a = torch.empty((n, n), dtype=torch.float64, device=dev)
    ee.mat_set_dev(n, a.data_ptr(), n, 2, 1)
    ee.sync()
    nrm = torch.linalg.norm(a).item()
    w_ref = torch.linalg.eigvalsh(a).cpu().numpy()               # cuSOLVER dsyevd
    w = torch.zeros(n, dtype=torch.float64, device=dev)
    z = torch.zeros(8, dtype=torch.float64, device=dev)
    work = a.clone()
    torch.cuda.synchronize()
    ee.eigen_s_dev(n, work.data_ptr(), n, w.data_ptr(), z.data_ptr(), n, nvec=0, mode="N")
    ee.sync()
    w = w.cpu().numpy()
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - w_ref).max() <= 10 * n * O.EPS * nrm


@pytest.mark.parametrize("n,mt", [(300, 0), (301, 0), (777, 3), (1000, 0)])
def test_prd_frank_entries_fixed_by_the_first_reflector(ee, n, mt):
    """See tests/test_oracle_golden.py::test_frank_band_form_is_not_unique_beyond_the_first_pair."""
    a = O.mat_set(n, mt)
    tol = 10 * n * O.EPS * np.linalg.norm(O.sym_from_upper(a))
    do, e1o, e2o = O.prd(F(a), 48)
    dg, e1g, e2g = ee.eigen_prd(n, F(a), 48)
    assert np.abs(dg[n - 2:] - do[n - 2:]).max() <= tol
    assert np.abs(e1g[n - 2:] - e1o[n - 2:]).max() <= tol
    assert abs(e2g[n - 1] - e2o[n - 1]) <= tol
    assert abs(e2g[n - 2]) <= tol and abs(e2o[n - 2]) <= tol


@pytest.mark.parametrize("mt", [7, 9, 10])
@pytest.mark.parametrize("solver", ["s", "sx"])
def test_helmert_families_7_9_10(ee, mt, solver):
    """benchmark/mat_set.f:651-729: A = H diag(w) H^T with the Frank spectrum (7), a Gaussian-like one (9) and the
    one read from benchmark/W.dat (10); w_test.f:142-156 gates."""
    n = 400
    a = O.mat_set(n, mt)
    w, z = np.zeros(n), np.zeros((n, n), order="F")
    (ee.eigen_sx if solver == "sx" else ee.eigen_s)(n, F(a), w, z)
    rel, ab = O.w_test(w, mt)
    assert ab < np.sqrt(O.EPS) * max(1.0, np.abs(w).max())
    res, orth = O.ev_test(O.sym_from_upper(a), w, z)
    assert res <= 10 and orth <= 10
