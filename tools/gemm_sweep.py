"""DMMA GEMM rate sweep on the shapes the path uses, next to cuBLAS (torch.matmul fp64) on the same box.
   EIGENEXA_B200_GEMM_CFG=<n> python tools/gemm_sweep.py"""
import sys, os, json
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
lib_stream = torch.cuda.ExternalStream(E.stream_ptr(), device=dev)
def timed(f, reps):
    f(); torch.cuda.synchronize(); E.sync()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(lib_stream):
            e0.record(); f(); e1.record()
        e1.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best
def rate(ta, tb, m, n, k, beta, reps=3):
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    A = torch.rand(ac, ar, dtype=torch.float64, device=dev); B = torch.rand(bc, br, dtype=torch.float64, device=dev)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    ours = timed(lambda: E.dgemm_dev(ta, tb, m, n, k, -1.0, A.data_ptr(), ar, B.data_ptr(), br, beta, Cm.data_ptr(), m), reps)
    # cuBLAS on the same operands: row-major torch tensors X hold X^T column-major, C^T = opB^T opA^T
    At = A if ta == "N" else A.T      # (k x m) = opA^T
    Bt = B if tb == "N" else B.T      # (n x k) = opB^T
    def cub():
        if beta == 0.0: torch.matmul(Bt, At, out=Cm)
        else: Cm.addmm_(Bt, At, beta=beta, alpha=-1.0)
    cu = timed(cub, reps)
    fl = 2.0 * m * n * k
    return {"ours_tflops": round(fl / ours / 1e12, 2), "cublas_tflops": round(fl / cu / 1e12, 2)}
out = {"cfg": os.environ.get("EIGENEXA_B200_GEMM_CFG", "default")}
for k in (96, 256, 512):
    out[f"NT_24576x24576x{k}_b1 (syr2k)"] = rate("N", "T", 24576, 24576, k, 1.0)
for k in (128, 256):
    out[f"NN_24576x24576x{k}_b1 (Z+=V*SS)"] = rate("N", "N", 24576, 24576, k, 1.0)
for m in (128, 256):
    out[f"TN_{m}x24576x24576_b0 (V^T Z)"] = rate("T", "N", m, 24576, 24576, 0.0)
out["NN_8192^3_b0"] = rate("N", "N", 8192, 8192, 8192, 0.0)
out["NN_12288x24576x12288_b0 (merge)"] = rate("N", "N", 12288, 24576, 12288, 0.0)
print(json.dumps(out))
E.eigen_free()
