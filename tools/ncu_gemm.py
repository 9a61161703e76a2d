"""Short ncu target for the DMMA GEMM: two launches per shape on the shapes of the path.
   python tools/ncu_gemm.py [small|all]"""
import sys, torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
def run(ta, tb, m, n, k, beta):
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    A = torch.rand(ac, ar, dtype=torch.float64, device=dev); B = torch.rand(bc, br, dtype=torch.float64, device=dev)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    for _ in range(2):
        E.dgemm_dev(ta, tb, m, n, k, -1.0, A.data_ptr(), ar, B.data_ptr(), br, beta, Cm.data_ptr(), m)
    E.sync()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
run("N", "T", 16384, 16384, 256, 1.0)         # trailing update, K = 2*128
if which == "all":
    run("N", "N", 8192, 8192, 8192, 0.0)      # merge GEMM
    run("T", "N", 256, 16384, 16384, 0.0)     # V^T Z
E.eigen_free()
print("ok")
