"""GPU parity: the DMMA GEMM against torch fp64 matmul (the one floating-point kernel that
keeps a plain torch reference; tolerance = K * eps * |A||B| elementwise bound)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


# pads (lda - rows, ldb - rows): even leading dimensions take the TMA kernel, odd ones the cp.async kernel
@pytest.mark.parametrize("pads", [(4, 2), (3, 1)])
@pytest.mark.parametrize("alpha,beta", [(-1.25, 0.5), (-1.0, 1.0), (1.0, 0.0)])
@pytest.mark.parametrize("ta,tb,m,n,k", [("N", "T", 300, 200, 96), ("T", "N", 128, 333, 1000), ("N", "N", 257, 129, 128),
                                         ("T", "T", 64, 64, 37), ("N", "N", 1, 1, 1), ("N", "T", 1000, 1000, 96),
                                         ("T", "N", 77, 1025, 2049), ("N", "N", 1031, 515, 5),
                                         # many tiles per resident CTA: the persistent loop wraps the stage ring
                                         ("N", "T", 4100, 3900, 256), ("N", "N", 3000, 5000, 700), ("T", "N", 256, 6000, 3000)])
def test_dgemm_matches_torch(ee, ta, tb, m, n, k, alpha, beta, pads):
    import torch
    g = torch.Generator(device="cpu").manual_seed(m * 7 + n * 3 + k)
    dev = torch.device("cuda:0")
    # column-major buffers: store X^T row-major
    def colmajor(rows, cols, ld):
        t = torch.zeros(cols, ld, dtype=torch.float64)
        t[:, :rows] = torch.rand(cols, rows, generator=g, dtype=torch.float64) - 0.5
        return t.to(dev)
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = ar + pads[0], br + pads[1], m + 2
    A, B, Cm = colmajor(ar, ac, lda), colmajor(br, bc, ldb), colmajor(m, n, ldc)
    C0 = Cm.clone()
    torch.cuda.synchronize()     # the library runs on its own stream: torch's copies must have landed
    ee.dgemm_dev(ta, tb, m, n, k, alpha, A.data_ptr(), lda, B.data_ptr(), ldb, beta, Cm.data_ptr(), ldc)
    ee.sync()
    Am = A[:, :ar].T
    Bm = B[:, :br].T
    opA = Am if ta == "N" else Am.T
    opB = Bm if tb == "N" else Bm.T
    ref = alpha * (opA @ opB) + beta * C0[:, :m].T
    got = Cm[:, :m].T
    tol = 4 * (k + 2) * 2.0 ** -52 * float((opA.abs() @ opB.abs()).max() + C0.abs().max())
    assert float((got - ref).abs().max()) <= tol
    # padding untouched
    assert torch.equal(Cm[:, m:], C0[:, m:])


@pytest.mark.parametrize("m,n,k,grid", [(300, 280, 96, (1, 1, 0, 0)), (5000, 5000, 96, (1, 1, 0, 0)), (9000, 9000, 256, (1, 1, 0, 0)),
                                        (4000, 2000, 256, (1, 2, 0, 1)), (2600, 5100, 96, (2, 1, 1, 0)),
                                        (3000, 1500, 256, (2, 4, 1, 3))])
def test_dgemm_staircase(ee, m, n, k, grid):
    """Trailing-update form (eigen_common_2update, src/eigen_t1.F:250-306): C -= A B^T on the tiles that reach
    the upper staircase of the cyclic local matrix; every element with global row <= global col must be
    updated, tiles entirely below the staircase must stay untouched (the driver keeps zeros there)."""
    import torch
    px, py, x, y = grid
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    A = (torch.rand(k, m, generator=g, dtype=torch.float64) - 0.5).to(dev)
    B = (torch.rand(k, n, generator=g, dtype=torch.float64) - 0.5).to(dev)
    C0 = (torch.rand(n, m, generator=g, dtype=torch.float64) - 0.5).to(dev)
    Cm = C0.clone()
    torch.cuda.synchronize()
    ee.dgemm_tri_dev("N", "T", m, n, k, -1.0, A.data_ptr(), m, B.data_ptr(), n, 1.0, Cm.data_ptr(), m, px, py, x, y)
    ee.sync()
    ref = C0 - B.T @ A                                    # (n x m) = C^T
    gr = (torch.arange(m, device=dev) * px + x)[None, :]
    gc = (torch.arange(n, device=dev) * py + y)[:, None]
    upper = gr <= gc
    tol = 4 * (k + 2) * 2.0 ** -52 * float((B.abs().T @ A.abs()).max() + 1.0)
    assert float(((Cm - ref).abs() * upper).max()) <= tol
    changed = (Cm != C0)
    # an element below the staircase may only change together with its whole tile reaching the staircase
    assert float(((Cm - ref).abs() * changed).max()) <= tol
    assert int((changed & upper).sum()) >= int(upper.sum()) - m - n     # (exact zeros of the update aside)
