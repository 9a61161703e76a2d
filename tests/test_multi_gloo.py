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


@pytest.mark.parametrize("px,py,order,n,i0,cur,iblk", [(1, 2, "C", 40, 17, 8, 1), (2, 2, "C", 53, 20, 16, 1), (2, 4, "C", 97, 33, 32, 1),
                                                       (2, 4, "R", 97, 34, 31, 2), (3, 2, "R", 61, 5, 12, 2), (2, 4, "C", 30, 1, 29, 1)])
def test_trbak_panel_pack_allgather_unpack_algebra(px, py, order, n, i0, cur, iblk):
    """Index algebra of the multi-rank V panel of the back-transformation (ee_trbak.cu: pack_v_kernel, all-gather over the
    world, unpack_v_kernel) restated in numpy: every rank packs the pieces it owns of the reflector columns i0 .. i0+cur-1
    (rows g = x mod px, columns gc = y mod py; reflector of column gc has gc-(iblk-1) rows), the gathered pieces must unpack
    to the replicated panel V(g, c) = A(g, i0+c) for g < i0+c-(iblk-1), 0 elsewhere (src/trbakwy4.F:686-733)."""
    rng = np.random.default_rng(n + i0)
    A = rng.standard_normal((n, n))
    P = px * py
    rows = i0 + cur - iblk
    cyc = lambda G, Pn, r: (G - r + Pn - 1) // Pn if G > r else 0
    nrl_max = (n + px - 1) // px
    cols_max, rows_max = cur // py + 2, (nrl_max + 2) // 2 * 2
    gathered = np.zeros((P, cols_max, rows_max))
    for x in range(px):
        for y in range(py):
            wr = x * py + y if order == "R" else x + y * px
            a_loc = A[x::px, y::py]
            lc0 = cyc(i0, py, y)
            for lc in range(cols_max):
                gc = (lc0 + lc) * py + y
                for jl in range(rows_max):
                    g = jl * px + x
                    if i0 <= gc < i0 + cur and g < gc - (iblk - 1) and g < rows:
                        gathered[wr, lc, jl] = a_loc[jl, lc0 + lc]
    V = np.zeros((n, cur))
    for c in range(cur):
        gc = i0 + c
        for g in range(rows):
            xo, yo = g % px, gc % py
            wr = xo * py + yo if order == "R" else xo + yo * px
            V[g, c] = gathered[wr, gc // py - cyc(i0, py, yo), g // px]
    want = np.zeros((n, cur))
    for c in range(cur):
        L = max(i0 + c - (iblk - 1), 0)
        want[:L, c] = A[:L, i0 + c]
    assert np.array_equal(V, want)


@pytest.mark.parametrize("px,py,n,nvec", [(1, 2, 23, 23), (2, 2, 31, 20), (2, 4, 41, 41), (4, 2, 37, 30)])
def test_distributed_ev_test_chunk_algebra(px, py, n, nvec):
    """K-chunk algebra of the on-grid ev_test (ee_matset.cu: ev_test_dist): R_loc = sum over chunks of
    [all-gather over y of KC/py local columns of A] x [all-gather over x of KC/px local rows of Z, permuted to the K order
    of the gathered A chunk]; G_loc = sum over row chunks of Z_loc^T [all-gather over y of Z rows], summed over x."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)); A = A + A.T
    Z = rng.standard_normal((n, nvec))
    w = rng.standard_normal(nvec)
    KC = 8
    while KC % px or KC % py:
        KC += 4
    kx, ky = KC // px, KC // py
    nvl_max = (nvec + py - 1) // py
    r2 = g2 = 0.0
    for x in range(px):
        for y in range(py):
            a_x = [A[x::px, yy::py] for yy in range(py)]       # what the ranks of my y group hold (same x)
            z_y = [Z[xx::px, y::py] for xx in range(px)]        # what the ranks of my x group hold (same y)
            nrl, nvl = a_x[0].shape[0], z_y[0].shape[1]
            R = np.zeros((nrl, nvl))
            for k0 in range(0, n, KC):
                Ag = np.zeros((nrl, KC))
                for yy in range(py):
                    blk = a_x[yy][:, k0 // py:k0 // py + ky]
                    Ag[:, yy * ky:yy * ky + blk.shape[1]] = blk
                Zg = np.zeros((px, kx, nvl))
                for xx in range(px):
                    blk = z_y[xx][k0 // px:k0 // px + kx, :]
                    Zg[xx, :blk.shape[0], :] = blk
                Zc = np.zeros((KC, nvl))
                for t in range(KC):
                    yp, cc = t // ky, t % ky
                    kk = cc * py + yp
                    Zc[t] = Zg[kk % px, kk // px]
                R += Ag @ Zc
            R -= z_y[x] * w[y::py][None, :]
            want = (A @ Z - Z * w[None, :])[x::px, y::py]
            assert np.allclose(R, want, atol=1e-11)
            r2 += (R ** 2).sum()
            if x == 0:
                # G_loc (nvl x py*nvl_max), summed over the x group
                G = np.zeros((nvl, py * nvl_max))
                for xx in range(px):
                    rows_all = np.zeros((Z[xx::px].shape[0], py * nvl_max))
                    for yy in range(py):
                        blk = Z[xx::px, yy::py]
                        rows_all[:, yy * nvl_max:yy * nvl_max + blk.shape[1]] = blk
                    G += Z[xx::px, y::py].T @ rows_all
                for i in range(nvl):
                    G[i, y * nvl_max + i] -= 1.0
                g2 += (G ** 2).sum()
    assert np.isclose(r2, ((A @ Z - Z * w[None, :]) ** 2).sum(), rtol=1e-10)
    assert np.isclose(g2, ((Z.T @ Z - np.eye(nvec)) ** 2).sum(), rtol=1e-10)
