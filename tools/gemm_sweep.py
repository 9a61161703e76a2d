"""DMMA GEMM rate sweep on the shapes the path uses.  python tools/gemm_sweep.py"""
import sys, time, json
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
def rate(ta, tb, m, n, k, beta, reps=4):
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    A = torch.rand(ac, ar, dtype=torch.float64, device=dev); B = torch.rand(bc, br, dtype=torch.float64, device=dev)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    f = lambda: E.dgemm_dev(ta, tb, m, n, k, -1.0, A.data_ptr(), ar, B.data_ptr(), br, beta, Cm.data_ptr(), m)
    f(); E.sync(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); f(); E.sync(); best = min(best, time.perf_counter() - t0)
    return round(2.0 * m * n * k / best / 1e12, 2)
out = {}
for k in (96, 128, 192, 256, 384, 512):
    out[f"NT_24576x24576x{k}_b1 (syr2k)"] = rate("N", "T", 24576, 24576, k, 1.0)
for k in (128, 256):
    out[f"NN_24576x24576x{k}_b1 (Z+=V*SS)"] = rate("N", "N", 24576, 24576, k, 1.0)
for m in (128, 256):
    out[f"TN_{m}x24576x24576_b0 (V^T Z)"] = rate("T", "N", m, 24576, 24576, 0.0)
out["NN_12288^3_b0 (merge)"] = rate("N", "N", 12288, 12288, 12288, 0.0)
print(json.dumps(out, indent=1))
E.eigen_free()
