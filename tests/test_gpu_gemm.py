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
