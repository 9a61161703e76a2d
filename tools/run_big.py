"""Device-resident eigen_s at large N with on-device generation and ev_test.  python tools/run_big.py N [N...]"""
import sys, time, json
import numpy as np
import torch
sys.path.insert(0, ".")
import eigenexa_b200 as E

dev = torch.device("cuda:0")
E.eigen_init(None, "C")
for n in [int(v) for v in sys.argv[1:]]:
    a = torch.empty((n, n), dtype=torch.float64, device=dev)   # column-major n x n (symmetric anyway)
    w = torch.empty(n, dtype=torch.float64, device=dev)
    z = torch.empty((n, n), dtype=torch.float64, device=dev)
    E.mat_set_dev(n, a.data_ptr(), n, 2, 1)
    a0 = a.clone()
    torch.cuda.synchronize()
    E.launch_count(True)
    t0 = time.perf_counter()
    E.eigen_s_dev(n, a.data_ptr(), n, w.data_ptr(), z.data_ptr(), n)
    t1 = time.perf_counter() - t0
    tm = E.last_timings()
    flops = 4.0 / 3.0 * n**3 + 2.0 * n**3
    r = {"n": n, "wall_s": t1, "trd_s": tm[1], "dc_s": tm[2], "trbak_s": tm[3], "tflops(trd+bak)": flops / t1 / 1e12,
         "dc_flops": tm[13], "dc_deflated": tm[14], "dc_defl_s": tm[17], "dc_perm_s": tm[18], "dc_sec_s": tm[19], "dc_gemm_s": tm[20], "dc_sort_s": tm[21],
         "launches": E.launch_count(True), "mem_peak_GB": torch.cuda.max_memory_allocated() / 1e9}
    print(json.dumps(r)); sys.stdout.flush()
    del a
    res, orth = E.ev_test_dev(n, n, a0.data_ptr(), n, w.data_ptr(), z.data_ptr(), n)
    print(json.dumps({"n": n, "residual": res, "orth": orth, "wmin": float(w[0]), "wmax": float(w[-1])})); sys.stdout.flush()
    del a0, w, z
    torch.cuda.empty_cache()
E.eigen_free()
