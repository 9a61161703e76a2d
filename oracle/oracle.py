"""CPU oracle for the eigen_s hot path -- TEST INFRASTRUCTURE ONLY.

Python face of ``oracle/eigenexa_oracle.c`` (see that file's header for the reference
file:line each routine restates).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product (``eigenexa_b200``) never does.

Third-party arithmetic that is *not* under /root/reference: the reference's tridiagonal
divide & conquer is ScaLAPACK PDSTEDC-derived code calling LAPACK DSTEDC/DLAED4 (no version
pinned by the reference, licence text says LAPACK-3.4.2 / ScaLAPACK-2.0.2).  The oracle
uses LAPACK ``dstevd`` (= DSTEDC) from SciPy's bundled OpenBLAS for that stage.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EPS = 2.0 ** -52  # get_constant_eps, src/eigen_libs0.F:2446-2459


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "eigenexa_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        for name in ("ora_loop_start", "ora_loop_end", "ora_translate_l2g", "ora_translate_g2l",
                     "ora_owner_node", "ora_owner_index"):
            f = getattr(L, name)
            f.argtypes = [C.c_int] * 3
            f.restype = C.c_int
        L.ora_grid_dims.argtypes = [C.c_int, ip, ip]
        L.ora_grid_coords.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char, ip, ip]
        L.ora_get_matdims.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char, C.c_int, ip, ip]
        L.ora_scaling.argtypes = [C.c_int, dp, C.c_int]
        L.ora_scaling.restype = C.c_double
        L.ora_trd.argtypes = [C.c_int, dp, C.c_int, dp, dp, C.c_int]
        L.ora_prd.argtypes = [C.c_int, dp, C.c_int, dp, dp, dp, C.c_int]
        L.ora_trbakwy.argtypes = [C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int]
        L.ora_bisect.argtypes = [C.c_int, dp, dp, dp]
        L.ora_mat_set_local.argtypes = [C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_uint64]
        L.ora_w_frank.argtypes = [C.c_int, dp]
        L.ora_rand_ij.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.ora_rand_ij.restype = C.c_double
        _LIB = L
    return _LIB


def _dp(x: np.ndarray):
    assert x.dtype == np.float64
    return x.ctypes.data_as(C.POINTER(C.c_double))


# ----------------------------------------------------------------------------------------
# index algebra / API helpers
# ----------------------------------------------------------------------------------------
def loop_start(i, nnod, inod): return lib().ora_loop_start(i, nnod, inod)
def loop_end(i, nnod, inod): return lib().ora_loop_end(i, nnod, inod)
def translate_l2g(i, nnod, inod): return lib().ora_translate_l2g(i, nnod, inod)
def translate_g2l(i, nnod, inod): return lib().ora_translate_g2l(i, nnod, inod)
def owner_node(i, nnod, inod): return lib().ora_owner_node(i, nnod, inod)
def owner_index(i, nnod, inod): return lib().ora_owner_index(i, nnod, inod)


def grid_dims(nnod: int):
    x, y = C.c_int(), C.c_int()
    lib().ora_grid_dims(nnod, C.byref(x), C.byref(y))
    return x.value, y.value


def grid_coords(inod: int, x_nnod: int, y_nnod: int, order: str = "C"):
    x, y = C.c_int(), C.c_int()
    lib().ora_grid_coords(inod, x_nnod, y_nnod, order.encode()[:1], C.byref(x), C.byref(y))
    return x.value, y.value


def get_matdims(n, x_nnod=1, y_nnod=1, m_f=48, m_b=128, mode="O", fs=True):
    nx, ny = C.c_int(), C.c_int()
    lib().ora_get_matdims(n, x_nnod, y_nnod, m_f, m_b, mode.encode()[:1], int(fs), C.byref(nx), C.byref(ny))
    return nx.value, ny.value


# ----------------------------------------------------------------------------------------
# stages
# ----------------------------------------------------------------------------------------
def scaling(a: np.ndarray) -> float:
    """In place on a Fortran-ordered n x n array; returns sigma."""
    assert a.flags.f_contiguous
    return lib().ora_scaling(a.shape[1], _dp(a), a.shape[0])


def trd(a: np.ndarray, m_f: int = 48):
    """Tridiagonalise (upper triangle of) Fortran-ordered ``a`` in place.

    Returns (d, e) with e[i] coupling rows i-1,i (0-based e[0] = 0)."""
    assert a.flags.f_contiguous
    n = a.shape[1]
    d = np.zeros(n)
    e = np.zeros(n)
    lib().ora_trd(n, _dp(a), a.shape[0], _dp(d), _dp(e), m_f)
    return d, e


def prd(a: np.ndarray, m_f: int = 48):
    """Reduce (upper triangle of) Fortran-ordered ``a`` to penta-diagonal form in place
    (eigen_prd, src/eigen_prd.F:341-580).  Returns (d, e1, e2): e1[i] = T(i-1,i), e2[i] = T(i-2,i)."""
    assert a.flags.f_contiguous
    n = a.shape[1]
    d, e1, e2 = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().ora_prd(n, _dp(a), a.shape[0], _dp(d), _dp(e1), _dp(e2), m_f)
    return d, e1, e2


def band_from(d: np.ndarray, e1: np.ndarray, e2: np.ndarray) -> np.ndarray:
    """Dense symmetric penta-diagonal matrix of (d, e1, e2)."""
    n = d.shape[0]
    b = np.diag(d)
    if n > 1:
        b += np.diag(e1[1:], 1) + np.diag(e1[1:], -1)
    if n > 2:
        b += np.diag(e2[2:], 2) + np.diag(e2[2:], -2)
    return b


def band_eig(d: np.ndarray, e1: np.ndarray, e2: np.ndarray):
    """Eigen-decomposition of the penta-diagonal matrix (stands in for eigen_dcx, src/dcx.F:75): LAPACK
    dsbevd from SciPy's OpenBLAS on the upper band storage."""
    from scipy.linalg import eig_banded
    n = d.shape[0]
    if n == 1:
        return d.copy(), np.ones((1, 1), order="F")
    ab = np.zeros((3, n))
    ab[2] = d
    ab[1, 1:] = e1[1:]
    ab[0, 2:] = e2[2:]
    w, z = eig_banded(ab, lower=False)
    return w, np.asfortranarray(z)


def trbakwy(a: np.ndarray, e: np.ndarray, z: np.ndarray, m_b: int = 128, nvec: int | None = None, iblk: int = 1):
    """Back-transform the first nvec columns of Fortran-ordered z in place (iblk = 1: reflectors of
    eigen_trd, length i-1; iblk = 2: reflectors of eigen_prd, length i-2, with e = e2)."""
    assert a.flags.f_contiguous and z.flags.f_contiguous
    n = a.shape[1]
    nvec = z.shape[1] if nvec is None else nvec
    beta = np.array(e, dtype=np.float64, copy=True)
    lib().ora_trbakwy(n, nvec, _dp(a), a.shape[0], _dp(z), z.shape[0], _dp(beta), m_b, iblk)
    return z


def bisect(d: np.ndarray, e: np.ndarray) -> np.ndarray:
    n = d.shape[0]
    w = np.zeros(n)
    lib().ora_bisect(n, _dp(np.ascontiguousarray(d)), _dp(np.ascontiguousarray(e)), _dp(w))
    return w


def tridiag_eig(d: np.ndarray, e: np.ndarray):
    """LAPACK divide & conquer (dstevd -> DSTEDC) on T=(d, e[1:])."""
    from scipy.linalg import lapack
    n = d.shape[0]
    if n == 1:
        return d.copy(), np.ones((1, 1), order="F")
    w, z, info = lapack.dstevd(d.copy(), e[1:].copy(), compute_v=1)
    if info != 0:
        raise RuntimeError(f"dstevd info={info}")
    return w, np.asfortranarray(z)


def eigen_s(a: np.ndarray, nvec: int | None = None, m_f: int = 48, m_b: int = 128, mode: str = "A"):
    """Restatement of eigen_s0 (src/eigen_s.F:30-305) on a 1x1 grid.

    ``a``: Fortran-ordered n x n, upper triangle read, destroyed.  Returns (w, z)."""
    assert a.flags.f_contiguous
    n = a.shape[1]
    nvec = n if nvec is None else nvec
    if nvec == 0:
        mode = "N"
    m_f = max(1, min(m_f, n))
    m_b = max(1, min(m_b, n))
    sigma = scaling(a)
    if np.isnan(sigma):
        return np.full(n, np.nan), None
    d, e = trd(a, m_f)
    if mode == "N":
        # NB: the reference leaves w un-rescaled in mode 'N' (goto 99999 at eigen_s.F:232-234)
        return bisect(d, e), None
    w, z = tridiag_eig(d, e)
    z = np.asfortranarray(z[:, :abs(nvec)])
    trbakwy(a, e, z, m_b)
    if sigma != 1.0 and sigma != 0.0:
        w = w * (1.0 / sigma)
    return w, z


def eigen_sx(a: np.ndarray, nvec: int | None = None, m_f: int = 48, m_b: int = 128, mode: str = "A"):
    """Restatement of eigen_sx (src/eigen_sx.F:30-308) on a 1x1 grid: scaling, eigen_prd, band
    eigensolver, back-transformation with MBAND = 2."""
    assert a.flags.f_contiguous
    n = a.shape[1]
    nvec = n if nvec is None else nvec
    if nvec == 0:
        mode = "N"
    m_b = max(1, min(m_b, n))
    sigma = scaling(a)
    if np.isnan(sigma):
        return np.full(n, np.nan), None
    d, e1, e2 = prd(a, m_f)
    w, z = band_eig(d, e1, e2)
    if mode == "N":
        return w, None
    z = np.asfortranarray(z[:, :abs(nvec)])
    trbakwy(a, e2, z, m_b, iblk=2)
    if sigma != 1.0 and sigma != 0.0:
        w = w * (1.0 / sigma)
    return w, z


# ----------------------------------------------------------------------------------------
# benchmark/mat_set.f restated
# ----------------------------------------------------------------------------------------
MAT_FRANK, MAT_TOEPLITZ, MAT_RANDOM, MAT_FRANK2 = 0, 1, 2, 3


def w_set(n: int, mtype: int, seed: int = 0) -> np.ndarray | None:
    """Prescribed spectra of benchmark/mat_set.f:606-729 (unsorted, as generated)."""
    i = np.arange(1, n + 1, dtype=np.float64)
    eps4 = np.sqrt(np.sqrt(EPS))
    if mtype in (0, 3, 7):
        w = np.zeros(n)
        lib().ora_w_frank(n, _dp(w))
        return w
    if mtype == 4:
        return i - 1.0
    if mtype == 5:
        return np.sin(np.pi * 5 * i / (n - 1) + eps4) ** 3
    if mtype == 6:
        ii = np.arange(1, n + 1)
        return (ii % 5 + ii % 2).astype(np.float64)
    if mtype == 8:
        return np.random.default_rng(seed).random(n)
    if mtype == 9:
        s = np.random.default_rng(seed).random(n)
        s = np.maximum(s, 1e-300)
        return np.sqrt(-2 * np.log(s)) * np.sin(2 * np.pi * s)
    if mtype == 10:
        # spectrum read from benchmark/W.dat (mat_set.f:714-729); its first 2000 values are committed as a fixture
        head = np.load(os.path.join(os.path.dirname(_HERE), "tests", "golden", "w_dat_head.npy"))
        if n > head.shape[0]:
            raise ValueError("mat_set type 10: only the first 2000 values of W.dat are committed")
        return head[:n].copy()
    return None


def helmert(n: int) -> np.ndarray:
    """Rows are the Helmert vectors h_i used by helmert_trans (mat_set.f:337-454)."""
    H = np.zeros((n, n))
    H[0, :] = 1.0 / np.sqrt(n)
    for i in range(2, n + 1):
        hi = np.sqrt(i - 1.0) * np.sqrt(float(i))
        H[i - 1, : i - 1] = 1.0 / hi
        H[i - 1, i - 1] = -(i - 1.0) / hi
    return H


def mat_set(n: int, mtype: int, seed: int = 1) -> np.ndarray:
    """Global test matrix (Fortran order).  Types 0-3 via the C generator, 4-10 Helmert."""
    if mtype in (0, 1, 2, 3):
        a = np.zeros((n, n), order="F")
        lib().ora_mat_set_local(mtype, n, _dp(a), n, 1, 1, 1, 1, seed)
        return a
    w = w_set(n, mtype)
    scale = max(1.0, np.abs(w).max())
    w_ = np.random.default_rng(0).permutation(w / scale)
    H = helmert(n)
    a = (H * w_[None, :]) @ H.T
    a = 0.5 * (a + a.T) * scale
    return np.asfortranarray(a)


def mat_set_local(n, mtype, x_nnod, y_nnod, x_inod, y_inod, lda=None, ncols=None, seed=1):
    """Local part of the 2D cyclic distribution (types 0-3)."""
    nr = loop_end(n, x_nnod, x_inod)
    nc = loop_end(n, y_nnod, y_inod)
    lda = max(nr, 1) if lda is None else lda
    ncols = max(nc, 1) if ncols is None else ncols
    a = np.zeros((lda, ncols), order="F")
    lib().ora_mat_set_local(mtype, n, _dp(a), lda, x_nnod, y_nnod, x_inod, y_inod, seed)
    return a


def scatter_cyclic(g: np.ndarray, x_nnod, y_nnod, x_inod, y_inod, lda=None, ncols=None):
    """global -> local (rows x_inod::x_nnod, cols y_inod::y_nnod), 1-based ids."""
    loc = g[x_inod - 1::x_nnod, y_inod - 1::y_nnod]
    lda = max(loc.shape[0], 1) if lda is None else lda
    ncols = max(loc.shape[1], 1) if ncols is None else ncols
    out = np.zeros((lda, ncols), order="F")
    out[:loc.shape[0], :loc.shape[1]] = loc
    return out


def gather_cyclic(parts: dict, n_rows: int, n_cols: int, x_nnod: int, y_nnod: int) -> np.ndarray:
    """parts[(x_inod,y_inod)] -> global n_rows x n_cols."""
    g = np.zeros((n_rows, n_cols), order="F")
    for (x, y), loc in parts.items():
        nr = len(range(x - 1, n_rows, x_nnod))
        nc = len(range(y - 1, n_cols, y_nnod))
        g[x - 1::x_nnod, y - 1::y_nnod] = loc[:nr, :nc]
    return g


# ----------------------------------------------------------------------------------------
# benchmark/ev_test.f and w_test.f metrics
# ----------------------------------------------------------------------------------------
def ev_test(a_full: np.ndarray, w: np.ndarray, z: np.ndarray):
    """(|AZ-ZW|_F/(N eps |A|_F), |Z^T Z - I|_F/(N eps)) -- ev_test.f:118-205.

    PASS gates of the reference: < 768 and < 8.  BASELINE.json's tighter gates: <= 10 each."""
    n = a_full.shape[0]
    nv = z.shape[1]
    r = a_full @ z - z * w[None, :nv]
    res = np.linalg.norm(r) / (n * EPS * np.linalg.norm(a_full))
    o = z.T @ z - np.eye(nv)
    orth = np.linalg.norm(o) / (n * EPS)
    return res, orth


def w_test(w: np.ndarray, mtype: int):
    """(max relative error, max absolute error) against the analytic spectrum, w_test.f:95-170."""
    ww = w_set(w.shape[0], mtype)
    if ww is None:
        return None
    ww = np.sort(ww)
    y = np.abs(w - ww)
    rel = np.where(ww == 0.0, 0.0, y / np.where(ww == 0.0, 1.0, np.abs(ww)))
    return rel.max(), y.max()


def sym_from_upper(a: np.ndarray) -> np.ndarray:
    u = np.triu(a)
    return np.asfortranarray(u + np.triu(a, 1).T)
