"""Generates tests/golden/*.npz: known answers the reference's own tests hold for this path.

  * Frank spectrum  w_k = 1/(2(1-cos(pi(2(n-k)+1)/(2n+1))))   benchmark/mat_set.f:638-647
  * eigen_get_matdims values quoted in SURVEY.md 8(a) (eigen_libs0.F:1254-1371 + CSTAB.F:73-131)
  * C/c_test.c:19-32  2x2 case -> (-3, -1)
  * benchmark/W.dat (spectrum of mat_set type 10), first 2000 values -> w_dat_head.npy
  * LAPACK dsytrd('U') (d, |e|) and dsyevd spectra of small seeded matrices (SciPy/OpenBLAS),
    the third-party arithmetic the reference calls but does not vendor.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from scipy.linalg import lapack

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
out = {}
for n in (10, 100, 1000):
    k = np.arange(1, n + 1)
    out[f"frank_w_{n}"] = 0.5 / (1.0 - np.cos(np.pi * (2 * (n - k) + 1) / (2 * n + 1)))
out["matdims"] = np.array([
    # n, px, py, nx, ny   (mode 'O', m_f 48, m_b 128, incl. the FS max)
    [10000, 1, 1, 10016, 10211],
    [50000, 2, 4, 25056, 12562],
    [1000, 1, 1, 1056, 1025],
])
out["ctest_w"] = np.array([-3.0, -1.0])
for n, mt in ((64, 2), (200, 0), (333, 2), (150, 3), (120, 1)):
    a = O.mat_set(n, mt)
    c, d, e, tau, info = lapack.dsytrd(np.array(a, order="F"), lower=0)
    w = np.linalg.eigvalsh(O.sym_from_upper(a))
    out[f"sytrd_d_{n}_{mt}"] = d
    out[f"sytrd_abs_e_{n}_{mt}"] = np.abs(e)
    out[f"eig_w_{n}_{mt}"] = w
    out[f"mat_{n}_{mt}_checksum"] = np.array([a.sum(), np.abs(a).max(), a[0, -1], a[n // 2, n // 3]])
# benchmark/W.dat: the prescribed spectrum of mat_set type 10 (mat_set.f:714-729); the first 2000 of its 100000 values
wdat = "/root/reference/benchmark/W.dat"
if os.path.exists(wdat):
    np.save(os.path.join(here, "w_dat_head.npy"), np.loadtxt(wdat, max_rows=2000))
np.savez_compressed(os.path.join(here, "golden.npz"), **out)
print("wrote", os.path.join(here, "golden.npz"), len(out), "arrays")
