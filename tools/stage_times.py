"""Stage / sub-stage seconds of one warm eigen_s (or eigen_sx) solve.  python tools/stage_times.py N [s|sx]"""
import sys, json, time
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E
n = int(sys.argv[1]); solver = sys.argv[2] if len(sys.argv) > 2 else "s"
prof = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
E.eigen_init(None, "C")
a0 = torch.empty((n, n), dtype=torch.float64, device=dev); a = torch.empty_like(a0)
w = torch.empty(n, dtype=torch.float64, device=dev); z = torch.empty((n, n), dtype=torch.float64, device=dev)
E.mat_set_dev(n, a0.data_ptr(), n, 2, 1)
f = E.eigen_sx_dev if solver == "sx" else E.eigen_s_dev
import os
print("persist env:", os.environ.get("EIGENEXA_B200_TRD_PERSIST"))
for rep in range(2):
    a.copy_(a0); torch.cuda.synchronize()
    E.set_profiling(prof if rep else 0)
    t0 = time.perf_counter()
    f(n, a.data_ptr(), n, w.data_ptr(), z.data_ptr(), n)
    wall = time.perf_counter() - t0
tm = E.last_timings()
res, orth = E.ev_test_dev(n, n, a0.data_ptr(), n, w.data_ptr(), z.data_ptr(), n)
names = {15: "symv_phase(persist)", 16: "p_phase(persist)", 31: "v_phase(persist)", 7: "pvec(prof2)", 8: "vvec(prof2)", 9: "trd_host_pre", 10: "trd_host_loop", 11: "trd_host_tail", 0: "h2d", 1: "trd", 2: "dc", 3: "trbak", 4: "d2h", 5: "symv", 6: "syr2k", 13: "dc_flops", 14: "dc_deflated",
         17: "dc_host_defl", 18: "dc_perm", 19: "dc_secular", 20: "dc_gemm", 21: "dc_sort"}
print(json.dumps({"n": n, "solver": solver, "wall": round(wall, 4), "residual": res, "orth": orth,
                  **{v: float(tm[k]) for k, v in names.items()}}))
E.eigen_free()
