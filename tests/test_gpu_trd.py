"""GPU parity: eigen_trd / eigen_trbakwy through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

F = lambda x: np.array(x, order="F", copy=True)


@pytest.mark.parametrize("n,mtype,mf", [(2, 0, 48), (3, 2, 48), (5, 2, 2), (7, 2, 1), (64, 0, 48), (100, 2, 7),
                                        (129, 2, 48), (300, 0, 48), (513, 2, 48), (1000, 0, 48), (1500, 2, 48),
                                        (777, 3, 32), (600, 1, 48)])
def test_trd_matches_oracle(ee, n, mtype, mf):
    a = O.mat_set(n, mtype)
    ao = F(a)
    do, eo = O.trd(ao, mf)
    ag = F(a)
    dg, eg = ee.eigen_trd(n, ag, mf)
    nrm = np.linalg.norm(O.sym_from_upper(a))
    tol = 10 * n * O.EPS * nrm  # BASELINE.json: (d, e) within 10 n eps |A|
    assert np.abs(dg - do).max() <= tol
    assert np.abs(eg - eo).max() <= tol
    # reflectors: column i (>=2) rows < i
    iu = np.triu_indices(n, 1)
    assert np.abs(ag[iu] - ao[iu]).max() <= 100 * tol / max(1.0, np.sqrt(nrm)) + tol


@pytest.mark.parametrize("n,mtype,mb,nvec", [(2, 0, 1, 2), (3, 2, 128, 3), (50, 2, 8, 50), (300, 0, 128, 300),
                                             (513, 2, 128, 513), (1000, 2, 128, 250), (700, 2, 64, 700)])
def test_trbak_matches_oracle(ee, n, mtype, mb, nvec):
    a = O.mat_set(n, mtype)
    ao = F(a)
    d, e = O.trd(ao, 48)
    w, zt = O.tridiag_eig(d, e)
    zt = F(zt[:, :nvec])
    zo = O.trbakwy(ao, e, F(zt), mb)
    zg = ee.eigen_trbakwy(n, ao, F(zt), e, mb, nvec=nvec)
    assert np.abs(zg - zo).max() <= 50 * n * O.EPS
    res, orth = O.ev_test(O.sym_from_upper(a), w, zg)
    assert res <= 10 and orth <= 10
