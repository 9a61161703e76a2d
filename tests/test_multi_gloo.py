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


@pytest.mark.parametrize("px,py,n,nvec", [(1, 2, 37, 37), (2, 2, 50, 41), (2, 4, 101, 101), (1, 8, 64, 50), (3, 2, 45, 45), (2, 4, 9, 9)])
def test_row_distributed_dc_exchange_algebra(px, py, n, nvec):
    """Index algebra of the row-distributed divide & conquer's last step (ee_dc.cu: pack_rows_kernel, grouped
    send/recv inside a grid row, unpack_rows_kernel), restated in numpy for grids the GPU tests cannot reach:
    world rank w = x + y*px owns the rows g = w (mod P) at local index g // P; rank (x, y) must end up with
    Z(jl, il) = Q(jl*px + x, ord[il*py + y]) -- the 2D cyclic layout of src/eigen_libs0.F:1986-2166."""
    rng = np.random.default_rng(px * 100 + py * 10 + n)
    P = px * py
    Q = rng.standard_normal((n, n))
    ordv = rng.permutation(n)
    cyc = lambda G, Pn, r: (G - r + Pn - 1) // Pn if G > r else 0
    # what every rank holds before the exchange
    loc = {w: Q[w::P, :] for w in range(P)}
    for x in range(px):
        # pack on every sender (x, ys): one buffer per destination y of the grid row, column-major (row fastest)
        send = {}
        for ys in range(py):
            w = x + ys * px
            nrow = loc[w].shape[0]
            assert nrow == cyc(n, P, w)
            for y in range(py):
                cols = [ordv[il * py + y] for il in range(cyc(nvec, py, y))]
                send[(ys, y)] = loc[w][:, cols]                       # (nrow, nvl_y)
        # unpack on every receiver (x, y)
        for y in range(py):
            nrl, nvl = cyc(n, px, x), cyc(nvec, py, y)
            z = np.zeros((nrl, nvl))
            for jl in range(nrl):
                ys, t = jl % py, jl // py
                z[jl, :] = send[(ys, y)][t, :]
            want = np.array([[Q[jl * px + x, ordv[il * py + y]] for il in range(nvl)] for jl in range(nrl)]).reshape(nrl, nvl)
            assert np.array_equal(z, want)
            # buffer sizes the driver checks: receive side fits into (nrow_loc + 4 rounded to 16) * n doubles
            nrow_loc = cyc(n, P, x + y * px)
            ldq = (max(nrow_loc, 1) + 4 + 15) // 16 * 16
            assert nrl * nvl <= ldq * n
