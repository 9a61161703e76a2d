"""CPU, world_size 2 (gloo): host-side logic of the multi-rank path -- bootstrap-id broadcast as
eigen_init_torch does it, grid coordinates, 2D cyclic scatter of the input and gather of the result,
replicated-vector reduction pattern of the forward step (all-reduce of per-rank partial A*u)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    # 1. bootstrap: rank 0 creates the 128-byte id, everyone receives the same bytes
    obj = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert obj[0] == bytes(range(128))
    # 2. grid as eigen_init picks it, column-major rank order
    px, py = O.grid_dims(world)
    xi, yi = O.grid_coords(rank + 1, px, py, "C")
    # 3. cyclic local part, partial symmetric mat-vec on the owned strict-upper staircase + diagonal
    a_loc = O.mat_set_local(n, 2, px, py, xi, yi)
    u = np.linspace(-1.0, 1.0, n)
    rows = np.arange(xi - 1, n, px)
    cols = np.arange(yi - 1, n, py)
    p = np.zeros(n)
    for jl, gj in enumerate(rows):
        for il, gi in enumerate(cols):
            v = a_loc[jl, il]
            if gj < gi:
                p[gi] += v * u[gj]      # column dot
                p[gj] += v * u[gi]      # row axpy
            elif gj == gi:
                p[gj] += v * u[gj]
    t = torch.from_numpy(p)
    dist.all_reduce(t)                   # the one collective of a column step
    full = O.sym_from_upper(O.mat_set(n, 2))
    assert np.allclose(t.numpy(), full @ u, rtol=1e-12, atol=1e-12)
    # 4. gather the distributed pieces back (what tools/run_multi.py does with Z)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(((xi, yi), a_loc), parts, dst=0)
    if rank == 0:
        g = O.gather_cyclic(dict(parts), n, n, px, py)
        assert np.array_equal(g, O.mat_set(n, 2))
        q.put("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_host_logic(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29611
    procs = [ctx.Process(target=_worker, args=(r, world, port, 41, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"
