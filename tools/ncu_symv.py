"""Short ncu target: first COLS columns of eigen_trd at size N (symv at L ~ N).  python tools/ncu_symv.py N COLS"""
import sys
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
n, cols = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
a = torch.empty((n, n), dtype=torch.float64, device=dev)
w = torch.empty(n, dtype=torch.float64, device=dev)
E.mat_set_dev(n, a.data_ptr(), n, 2, 1)
E.set_debug_maxcols(cols)
import ctypes
d = torch.empty(n, dtype=torch.float64, device=dev)
# stage-level call on device data is not exported; use the host stage entry on a small wrapper: eigen_s_dev mode N
# would run bisection on garbage, so call trd through eigen_s_dev with nvec=0 and ignore the values.
E.eigen_s_dev(n, a.data_ptr(), n, w.data_ptr(), 0, n, nvec=0, mode="N")
print("done", E.launch_count())
E.eigen_free()
