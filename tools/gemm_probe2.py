"""Small-K diagnosis: beta 0 vs 1, K = 128..1024, TMA kernel configs vs the cp.async kernel."""
import sys, os, json
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
lib_stream = torch.cuda.ExternalStream(E.stream_ptr(), device=dev)
def timed(f, reps=3):
    f(); torch.cuda.synchronize(); E.sync()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(lib_stream):
            e0.record(); f(); e1.record()
        e1.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best
out = {"cfg": os.environ.get("EIGENEXA_B200_GEMM_CFG", "default")}
m = n = 16384
for k in (128, 256, 512, 1024):
    A = torch.rand(k, m, dtype=torch.float64, device=dev); B = torch.rand(k, n, dtype=torch.float64, device=dev)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    for beta in (0.0, 1.0):
        t = timed(lambda: E.dgemm_dev("N", "T", m, n, k, -1.0, A.data_ptr(), m, B.data_ptr(), n, beta, Cm.data_ptr(), m))
        out[f"NT_k{k}_b{int(beta)}"] = round(2.0 * m * n * k / t / 1e12, 2)
print(json.dumps(out))
E.eigen_free()
