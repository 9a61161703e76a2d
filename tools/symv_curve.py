"""SYMV bandwidth vs trailing size L (device-resident trd with async event timing).
python tools/symv_curve.py N"""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
n = int(sys.argv[1])
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
a = torch.empty((n, n), dtype=torch.float64, device=dev)
w = torch.empty(n, dtype=torch.float64, device=dev)
E.mat_set_dev(n, a.data_ptr(), n, 2, 1)
a0 = a.clone()
for rep in range(2):
    a.copy_(a0); torch.cuda.synchronize()
    E.set_profiling(1)
    E.eigen_s_dev(n, a.data_ptr(), n, w.data_ptr(), 0, n, nvec=0, mode="N")
tr = E.symv_trace()
tm = E.last_timings()
L = n - 1 - np.arange(len(tr))
byts = 8.0 * L * (L - 1) / 2
pts = [n - 1, int(n * 0.9), int(n * 0.75), n // 2, n // 4, n // 8, 2000, 1000, 500]
rows = []
for p in pts:
    sel = (L <= p) & (L > p - 48)
    if sel.any():
        rows.append({"L": p, "GBs": float(byts[sel].sum() / (tr[sel].sum() * 1e-3) / 1e9), "us": float(tr[sel].mean() * 1e3)})
print(json.dumps({"n": n, "trd_s": tm[1], "symv_s": float(tr.sum() * 1e-3), "avg_GBs": float(byts.sum() / (tr.sum() * 1e-3) / 1e9),
                  "syr2k_s": tm[6], "syr2k_tflops": 2 / 3 * n**3 / tm[6] / 1e12, "curve": rows}))
E.eigen_free()
