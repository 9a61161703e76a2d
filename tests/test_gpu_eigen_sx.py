"""GPU parity of the penta-diagonal path (eigen_sx): eigen_prd, eigen_dcx, eigen_bisect2, the
back-transformation with nb = 2 and the full driver, through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

F = lambda x: np.array(x, order="F", copy=True)


@pytest.mark.parametrize("n,mtype,mf", [(4, 2, 48), (5, 2, 2), (6, 0, 48), (7, 2, 4), (64, 0, 48), (100, 2, 6),
                                        (129, 2, 48), (300, 0, 48), (513, 2, 48), (1000, 0, 48), (1500, 2, 48),
                                        (777, 3, 32), (600, 1, 48)])
def test_prd_matches_oracle(ee, n, mtype, mf):
    a = O.mat_set(n, mtype)
    ao = F(a)
    do, e1o, e2o = O.prd(ao, mf)
    ag = F(a)
    dg, e1g, e2g = ee.eigen_prd(n, ag, mf)
    full = O.sym_from_upper(a)
    nrm = np.linalg.norm(full)
    tol = 10 * n * O.EPS * nrm   # same bound BASELINE.json states for (d, e) of eigen_trd
    if mtype in (0, 3):
        # Frank matrices: columns i and i-1 coincide above the band, so the second reflector of the very
        # first pair is built from a vector that is exactly zero in exact arithmetic (mask(1) of
        # src/eigen_prd_t4x.F:204-208) and pure rounding noise otherwise: the band matrix is not unique,
        # only its spectrum (checked below) and the band structure are.
        assert e1g[0] == 0 and e2g[0] == 0 and e2g[1] == 0
    else:
        assert np.abs(dg - do).max() <= tol
        assert np.abs(e1g - e1o).max() <= tol
        assert np.abs(e2g - e2o).max() <= tol
    # the band matrix is orthogonally similar to A
    wb = np.linalg.eigvalsh(O.band_from(dg, e1g, e2g))
    assert np.abs(wb - np.linalg.eigvalsh(full)).max() <= tol


@pytest.mark.parametrize("n,kind", [(3, "rand"), (5, "rand"), (33, "rand"), (64, "rand"), (100, "rand"), (257, "rand"),
                                    (1000, "rand"), (500, "const"), (400, "graded"), (300, "zeroe"), (600, "prd")])
def test_dcx_penta(ee, n, kind):
    rng = np.random.default_rng(n)
    if kind == "rand":
        d, e1, e2 = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    elif kind == "const":
        d, e1, e2 = np.full(n, 6.0), np.full(n, -4.0), np.full(n, 1.0)     # biharmonic stencil: heavy deflation
    elif kind == "graded":
        g = 10.0 ** (-np.arange(n) / 40.0)
        d, e1, e2 = g.copy(), 0.1 * g, 0.01 * g
    elif kind == "zeroe":
        d, e1, e2 = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
        e1[::7] = 0.0; e2[::5] = 0.0
    else:
        d, e1, e2 = O.prd(O.mat_set(n, 2), 48)
    e1[0] = 0.0; e2[:2] = 0.0
    z = np.zeros((n, n), order="F")
    w = ee.eigen_dcx(n, d, e1, e2, z)
    B = O.band_from(d, e1, e2)
    wl = np.linalg.eigvalsh(B)
    nrm = max(np.linalg.norm(B), 1e-300)
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - wl).max() <= 10 * n * O.EPS * nrm
    res, orth = O.ev_test(B, w, z)
    assert res <= 10 and orth <= 10, (res, orth)


@pytest.mark.parametrize("n", [3, 50, 1000])
def test_bisect2(ee, n):
    rng = np.random.default_rng(n)
    d, e1, e2 = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    e1[0] = 0.0; e2[:2] = 0.0
    w = ee.eigen_bisect2(n, d, e1, e2)
    B = O.band_from(d, e1, e2)
    assert np.abs(w - np.linalg.eigvalsh(B)).max() <= 10 * n * O.EPS * max(np.linalg.norm(B), 1.0)


@pytest.mark.parametrize("n,kind", [(200, "chains"), (301, "chains"), (257, "integer"), (400, "biharmonic"), (300, "tridiag"),
                                    (64, "zero")])
def test_bisect2_exact_pivot_hits(ee, n, kind):
    """Band matrices whose leading principal minors are singular at (or next to) eigenvalues of the whole matrix: the
    unpivoted band L D L^T recurrence breaks down there; the 2x2-pivot safeguard (src/bisect2.F:393-678) must not."""
    rng = np.random.default_rng(n)
    if kind == "chains":        # e1 = 0: two interleaved tridiagonal chains, every eigenvalue (nearly) double
        d, e1, e2 = np.zeros(n), np.zeros(n), np.ones(n)
    elif kind == "integer":
        d, e1, e2 = (rng.integers(-2, 3, n).astype(float) for _ in range(3))
    elif kind == "biharmonic":
        d, e1, e2 = np.full(n, 6.0), np.full(n, -4.0), np.ones(n)
    elif kind == "tridiag":
        d, e1, e2 = np.zeros(n), np.ones(n), np.zeros(n)
    else:
        d, e1, e2 = np.zeros(n), np.zeros(n), np.zeros(n)
    e1[0] = 0.0; e2[:2] = 0.0
    w = ee.eigen_bisect2(n, d, e1, e2)
    B = O.band_from(d, e1, e2)
    assert np.all(np.isfinite(w)) and np.all(np.diff(w) >= 0)
    assert np.abs(w - np.linalg.eigvalsh(B)).max() <= 10 * n * O.EPS * max(np.linalg.norm(B), 1.0)


@pytest.mark.parametrize("n,mtype,mb,nvec", [(5, 2, 128, 5), (50, 2, 8, 50), (300, 0, 128, 300), (513, 2, 128, 513),
                                             (1000, 2, 128, 250)])
def test_trbak_nb2_matches_oracle(ee, n, mtype, mb, nvec):
    a = O.mat_set(n, mtype)
    ao = F(a)
    d, e1, e2 = O.prd(ao, 48)
    w, zt = O.band_eig(d, e1, e2)
    zt = F(zt[:, :nvec])
    zo = O.trbakwy(ao, e2, F(zt), mb, iblk=2)
    zg = ee.eigen_trbakwy(n, ao, F(zt), e2, mb, nvec=nvec, nb=2)
    assert np.abs(zg - zo).max() <= 50 * n * O.EPS
    res, orth = O.ev_test(O.sym_from_upper(a), w[:nvec], zg)
    assert res <= 10 and orth <= 10


@pytest.mark.parametrize("n,mtype,mf,mb", [(1, 0, 48, 128), (2, 2, 48, 128), (3, 2, 48, 128), (4, 2, 48, 128),
                                           (10, 2, 4, 4), (100, 0, 48, 128), (257, 2, 48, 128), (1000, 0, 48, 128),
                                           (1501, 2, 48, 128), (2000, 2, 48, 128), (900, 3, 32, 64)])
def test_eigen_sx_all_pairs(ee, n, mtype, mf, mb):
    a = O.mat_set(n, mtype)
    full = O.sym_from_upper(a)
    w, z = np.zeros(n), np.zeros((n, n), order="F")
    ee.eigen_sx(n, F(a), w, z, m_forward=mf, m_backward=mb, mode="A")
    wl = np.linalg.eigvalsh(full)
    tol = 10 * n * O.EPS * np.linalg.norm(full)
    assert np.abs(w - wl).max() <= tol
    wo, _ = O.eigen_sx(F(a), m_f=mf, m_b=mb)
    assert np.abs(w - wo).max() <= tol
    res, orth = O.ev_test(full, w, z)
    assert res <= 10 and orth <= 10, (res, orth)


def test_eigen_sx_mode_n_and_partial(ee):
    n = 700
    a = O.mat_set(n, 2)
    full = O.sym_from_upper(a)
    wl = np.linalg.eigvalsh(full)
    tol = 10 * n * O.EPS * np.linalg.norm(full)
    w, z = np.zeros(n), np.zeros((n, n), order="F")
    ee.eigen_sx(n, F(a), w, z, nvec=0, mode="N")
    assert np.abs(w - wl).max() <= tol
    nvec = 100
    w, z = np.zeros(n), np.zeros((n, nvec), order="F")
    ee.eigen_sx(n, F(a), w, z, nvec=nvec, mode="A")
    assert np.abs(w - wl).max() <= tol
    res, orth = O.ev_test(full, w[:nvec], z)
    assert res <= 10 and orth <= 10


def test_eigen_sx_edge_cases(ee):
    """Same conventions as eigen_s (src/eigen_sx.F:100-131,150-160,284-300): NaN input -> w = NaN, n <= 0 silent,
    lower triangle never read, odd m_forward rounded to the even width the pair loop needs (manual 4.4),
    mode 'X' (D&C then bisection refinement), scaling of huge matrices, a(1:3,1) bookkeeping."""
    n = 60
    a = O.mat_set(n, 2)
    full = O.sym_from_upper(a)
    wl = np.linalg.eigvalsh(full)
    tol = 10 * n * O.EPS * np.linalg.norm(full)
    w, z = np.zeros(n), np.zeros((n, n), order="F")
    bad = F(a); bad[3, 7] = np.nan
    ee.eigen_sx(n, bad, w, z)
    assert np.all(np.isnan(w))
    w[:] = 7.0
    ee.eigen_sx(0, F(a), w, z)
    assert np.all(w == 7.0)
    low = F(a); low[7, 3] = np.nan
    ee.eigen_sx(n, low, w, z)
    assert np.abs(w - wl).max() <= tol
    for mf in (1, 3, 7, 47):
        ag = F(a)
        ee.eigen_sx(n, ag, w, z, m_forward=mf, m_backward=5)
        assert np.abs(w - wl).max() <= tol
        res, orth = O.ev_test(full, w, z)
        assert res <= 10 and orth <= 10
        assert ag[0, 0] > 0 and ag[1, 0] > 0 and ag[2, 0] == -1.0
    ee.eigen_sx(n, F(a), w, z, mode="X")
    assert np.abs(w - wl).max() <= tol
    s = 1e200
    ee.eigen_sx(n, F(a * s), w, z)
    assert np.abs(w / s - wl).max() <= tol
    res, orth = O.ev_test(full, w / s, z)
    assert res <= 10 and orth <= 10


@pytest.mark.parametrize("n", [2000, 3001])
def test_eigen_sx_spectrum_equals_eigen_s(ee, n):
    """Both drivers must agree to the eigenvalue bound on the same matrix (different reductions, same spectrum)."""
    a = O.mat_set(n, 2)
    full = O.sym_from_upper(a)
    tol = 10 * n * O.EPS * np.linalg.norm(full)
    w1, z1 = np.zeros(n), np.zeros((n, n), order="F")
    w2, z2 = np.zeros(n), np.zeros((n, n), order="F")
    ee.eigen_s(n, F(a), w1, z1)
    ee.eigen_sx(n, F(a), w2, z2)
    assert np.abs(w1 - w2).max() <= tol
    for (w, z) in ((w1, z1), (w2, z2)):
        res, orth = O.ev_test(full, w, z)
        assert res <= 10 and orth <= 10


@pytest.mark.parametrize("solver", ["s", "sx"])
@pytest.mark.parametrize("kind", ["zero", "identity", "diag", "repeated_blocks", "rank_one", "arrow"])
def test_degenerate_matrices(ee, solver, kind):
    """sigma = 0 columns (g = 0, u = 0, beta = 1: src/eigen_trd_t2.F:574-583, src/eigen_prd_t4x.F:204-215), total
    deflation in the divide & conquer, multiple eigenvalues: no NaN, spectrum and eigenvector gates hold."""
    n = 130
    rng = np.random.default_rng(7)
    if kind == "zero":
        full = np.zeros((n, n))
    elif kind == "identity":
        full = np.eye(n)
    elif kind == "diag":
        full = np.diag(np.repeat(np.arange(1.0, 14.0), 10))
    elif kind == "repeated_blocks":
        b = rng.standard_normal((10, 10)); b = b + b.T
        full = np.kron(np.eye(13), b)
    elif kind == "rank_one":
        v = rng.standard_normal(n)
        full = np.outer(v, v)
    else:
        full = np.diag(np.arange(1.0, n + 1.0))
        full[:, -1] = 1.0; full[-1, :] = 1.0; full[-1, -1] = 5.0
    a = np.asfortranarray(np.triu(full))
    w, z = np.zeros(n), np.zeros((n, n), order="F")
    (ee.eigen_sx if solver == "sx" else ee.eigen_s)(n, a, w, z)
    assert np.all(np.isfinite(w)) and np.all(np.isfinite(z))
    nrm = max(np.linalg.norm(full), 1.0)
    assert np.abs(w - np.linalg.eigvalsh(full)).max() <= 10 * n * O.EPS * nrm
    r = full @ z - z * w[None, :]
    assert np.linalg.norm(r) <= 10 * n * O.EPS * nrm
    assert np.linalg.norm(z.T @ z - np.eye(n)) <= 10 * n * O.EPS
