"""Short ncu target: the first COLS columns of eigen_trd (solver s: symv_kernel) or eigen_prd (solver sx:
symv2_kernel) at size N, i.e. SYMV launches at L ~ N.   python tools/ncu_symv.py N COLS [s|sx]"""
import sys
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
n, cols = int(sys.argv[1]), int(sys.argv[2])
solver = sys.argv[3] if len(sys.argv) > 3 else "s"
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
a = torch.empty((n, n), dtype=torch.float64, device=dev)
w = torch.empty(n, dtype=torch.float64, device=dev)
E.mat_set_dev(n, a.data_ptr(), n, 2, 1)
E.set_debug_maxcols(cols)
# mode 'N' (nvec = 0): forward reduction + bisection; with debug_maxcols the reduction stops after COLS columns
f = E.eigen_sx_dev if solver == "sx" else E.eigen_s_dev
f(n, a.data_ptr(), n, w.data_ptr(), 0, n, nvec=0, mode="N")
print("done", E.launch_count())
E.eigen_free()
