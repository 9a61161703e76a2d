"""Per-class device time of the back-transformation (profiling level 2).  python tools/trbak_breakdown.py N"""
import sys, json, time
import numpy as np, torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
n = int(sys.argv[1]); dev = torch.device("cuda:0")
E.eigen_init(None, "C")
a = torch.empty((n, n), dtype=torch.float64, device=dev); w = torch.empty(n, dtype=torch.float64, device=dev)
z = torch.empty((n, n), dtype=torch.float64, device=dev)
E.mat_set_dev(n, a.data_ptr(), n, 2, 1)
E.set_debug_maxcols(2)          # skip (almost all of) the forward reduction: timings only
for rep in range(2):
    E.set_profiling(2 if rep else 0)
    t0 = time.perf_counter()
    E.eigen_s_dev(n, a.data_ptr(), n, w.data_ptr(), z.data_ptr(), n)
    t1 = time.perf_counter() - t0
tm = E.last_timings()
fl = 2.0 * n**3
print(json.dumps({"n": n, "wall": t1, "trbak_s": tm[3], "gatherV": tm[22], "S_Tinv": tm[23], "SS_gemm": tm[24], "T_SS": tm[25],
                  "update_gemm": tm[26], "host_alloc": tm[27], "host_loop": tm[28], "host_sync": tm[29], "host_free": tm[30], "SS_tflops": fl / 2 / tm[24] / 1e12, "update_tflops": fl / 2 / tm[26] / 1e12}))
E.eigen_free()
