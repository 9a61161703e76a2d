"""Multi-rank parity run (one process per GPU):  torchrun --nproc-per-node P tools/run_multi.py N [mtype] [mode] [check|nocheck] [s|sx] [C|R]
Each rank builds its 2D-cyclic part, calls eigen_s through the C ABI with host arrays, rank 0 gathers
w, Z and checks them against the oracle / ev_test; the distributed on-device ev_test (benchmark/ev_test.f on the
grid) is run on the same result and must agree with the host one.  Prints one JSON line on rank 0."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenexa_b200 as E
from oracle import oracle as O

n = int(sys.argv[1]); mtype = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "A"
check = (len(sys.argv) <= 4) or sys.argv[4] != "nocheck"
solver = sys.argv[5] if len(sys.argv) > 5 else "s"
order = sys.argv[6] if len(sys.argv) > 6 else "C"
solve = E.eigen_sx if solver == "sx" else E.eigen_s
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
E.eigen_init_torch(order)
nnod, px, py = E.eigen_get_procs(); inod, xi, yi = E.eigen_get_id()
assert nnod == world, E.last_error()
nx, ny = E.eigen_get_matdims(n)
a = O.mat_set_local(n, mtype, px, py, xi, yi, lda=nx, ncols=ny)
a_in = a.copy(order="F")
w = np.zeros(n); z = np.zeros((nx, ny), order="F")
dist.barrier(); t0 = time.perf_counter()
solve(n, a, w, z, mode=mode)
dist.barrier(); t1 = time.perf_counter() - t0
tm = E.last_timings()
if not check:
    if rank == 0:
        print(json.dumps({"n": n, "grid": f"{px}x{py}", "mode": mode, "seconds": t1, "h2d": tm[0], "trd": tm[1], "dc": tm[2],
                          "trbak": tm[3], "d2h": tm[4], "dc_gemm": tm[20], "dc_sec": tm[19]}), flush=True)
    E.eigen_free(); dist.destroy_process_group(); sys.exit(0)
dev_check = None
if mode != "N":
    # benchmark/ev_test.f on the device, distributed over the grid (collective: every rank calls it)
    dv = torch.device("cuda", lr)
    a_t = torch.from_numpy(np.ascontiguousarray(a_in.T)).to(dv)      # column-major lda x ny
    z_t = torch.from_numpy(np.ascontiguousarray(z.T)).to(dv)
    w_t = torch.from_numpy(w).to(dv)
    dev_check = E.ev_test_dev(n, n, a_t.data_ptr(), nx, w_t.data_ptr(), z_t.data_ptr(), nx)
parts = [None] * world if rank == 0 else None
nr, nc = E.eigen_loop_end(n, px, xi), E.eigen_loop_end(n, py, yi)
dist.gather_object(((xi, yi), z[:nr, :nc].copy(), w.copy()), parts, dst=0)
if rank == 0:
    ws = [p[2] for p in parts]
    assert all(np.array_equal(ws[0], v) for v in ws), "w differs between ranks"
    full = O.sym_from_upper(O.mat_set(n, mtype))
    wl = np.linalg.eigvalsh(full)
    out = {"n": n, "grid": f"{px}x{py}", "order": order, "solver": solver, "mode": mode, "seconds": t1, "trd": tm[1], "dc": tm[2], "trbak": tm[3],
           "w_err_over_tol": float(np.abs(w - wl).max() / (10 * n * O.EPS * np.linalg.norm(full)))}
    if mode != "N":
        Z = O.gather_cyclic({p[0]: p[1] for p in parts}, n, n, px, py)
        res, orth = O.ev_test(full, w, Z)
        out.update({"residual": res, "orth": orth, "residual_dev": dev_check[0], "orth_dev": dev_check[1]})
    print(json.dumps(out), flush=True)
E.eigen_free()
dist.destroy_process_group()
