"""GPU probe: FP64 peaks (cuBLAS DGEMM), our DMMA GEMM rates, eigen_trd stage timings.
Run on the GPU box: python tools/probe.py [n ...]"""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E

dev = torch.device("cuda:0")
out = {}

def time_cuda(fn, reps=5):
    fn(); torch.cuda.synchronize(); E.sync()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); E.sync(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best

# cuBLAS DGEMM
for N in (4096, 8192):
    a = torch.rand(N, N, dtype=torch.float64, device=dev); b = torch.rand(N, N, dtype=torch.float64, device=dev)
    t = time_cuda(lambda: torch.matmul(a, b))
    out[f"cublas_dgemm_{N}_tflops"] = 2 * N**3 / t / 1e12
    del a, b
# HBM read bandwidth (torch sum) and copy
x = torch.rand(1 << 29, dtype=torch.float64, device=dev)
t = time_cuda(lambda: x.sum()); out["read_sum_GBs"] = x.numel() * 8 / t / 1e9
y = torch.empty_like(x)
t = time_cuda(lambda: y.copy_(x)); out["copy_GBs"] = 2 * x.numel() * 8 / t / 1e9
del x, y
E.eigen_init(None, "C")
def gemm_rate(ta, tb, m, n, k, beta):
    ar, ac = (m, k) if ta == "N" else (k, m)
    br, bc = (k, n) if tb == "N" else (n, k)
    A = torch.rand(ac, ar, dtype=torch.float64, device=dev); B = torch.rand(bc, br, dtype=torch.float64, device=dev)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    t = time_cuda(lambda: E.dgemm_dev(ta, tb, m, n, k, 1.0, A.data_ptr(), ar, B.data_ptr(), br, beta, Cm.data_ptr(), m))
    return 2.0 * m * n * k / t / 1e12
out["dmma_NT_16384x16384x96_b1"] = gemm_rate("N", "T", 16384, 16384, 96, 1.0)
out["dmma_NN_16384x16384x128_b1"] = gemm_rate("N", "N", 16384, 16384, 128, 1.0)
out["dmma_TN_128x16384x16384_b0"] = gemm_rate("T", "N", 128, 16384, 16384, 0.0)
out["dmma_NN_8192_b0"] = gemm_rate("N", "N", 8192, 8192, 8192, 0.0)
out["dmma_TN_8192_b0"] = gemm_rate("T", "N", 8192, 8192, 8192, 0.0)
print(json.dumps(out, indent=1)); sys.stdout.flush()
ns = [int(v) for v in sys.argv[1:]] or [4000]
for n in ns:
    a = np.zeros((n, n), order="F")
    E.mat_set_host(n, a, 2, 1)
    E.set_profiling(0)
    t0 = time.perf_counter(); d, e = E.eigen_trd(n, a.copy(order="F")); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); d, e = E.eigen_trd(n, a.copy(order="F")); t1 = time.perf_counter() - t0
    E.set_profiling(2)
    E.eigen_trd(n, a.copy(order="F"))
    tm = E.last_timings()
    symv_bytes = 8.0 * sum((i * (i - 1)) / 2 for i in range(2, n))
    r = {"n": n, "trd_wall_s": t1, "symv_s(profiled)": tm[5], "symv_GBs": symv_bytes / tm[5] / 1e9,
         "syr2k_s": tm[6], "syr2k_tflops": (2.0 / 3.0) * n**3 / max(tm[6], 1e-9) / 1e12, "pvec_s": tm[7], "vvec_s": tm[8],
         "host_alloc_s": tm[9], "host_loop_s": tm[10], "host_tail_s": tm[11], "host_free_s": tm[12], "launches": E.launch_count(True)}
    print(json.dumps(r)); sys.stdout.flush()
E.eigen_free()

E.eigen_init(None, "C")
for n in ns:
    a = np.zeros((n, n), order="F"); E.mat_set_host(n, a, 2, 1)
    w = np.zeros(n); z = np.zeros((n, n), order="F")
    E.set_profiling(0)
    E.eigen_s(n, a.copy(order="F"), w, z)
    t0 = time.perf_counter(); E.eigen_s(n, a.copy(order="F"), w, z); t1 = time.perf_counter() - t0
    tm = E.last_timings()
    print(json.dumps({"n": n, "eigen_s_wall_s": t1, "h2d": tm[0], "trd": tm[1], "dc": tm[2], "trbak": tm[3], "d2h": tm[4]}))
E.eigen_free()
