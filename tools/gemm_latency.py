"""Host and device cost of one small dgemm call.  EIGENEXA_B200_GEMM_CFG=<n> python tools/gemm_latency.py"""
import sys, os, time, json
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
out = {"cfg": os.environ.get("EIGENEXA_B200_GEMM_CFG", "default")}
for (m, n, k) in ((64, 64, 64), (512, 512, 256), (2048, 2048, 96), (10000, 10000, 96)):
    A = torch.rand(k, m, dtype=torch.float64, device=dev); B = torch.rand(k, n, dtype=torch.float64, device=dev)
    C = torch.zeros(n, m, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    f = lambda: E.dgemm_dev("N", "T", m, n, k, -1.0, A.data_ptr(), m, B.data_ptr(), n, 1.0, C.data_ptr(), m)
    for _ in range(10): f()
    E.sync()
    reps = 200
    t0 = time.perf_counter()
    for _ in range(reps): f()
    t_host = (time.perf_counter() - t0) / reps
    E.sync()
    t_all = (time.perf_counter() - t0) / reps
    out[f"{m}x{n}x{k}"] = {"host_enqueue_us": round(t_host * 1e6, 1), "per_call_us_incl_gpu": round(t_all * 1e6, 1)}
print(json.dumps(out))
E.eigen_free()
