"""CPU model of the partial-sum scheme of the persistent panel kernel (ee_trd.cu: symv_strip, fold_triangle_g, strip_rows,
ntile_rows, the item shape (sw, sv), and the reader loops of the p phase), restated in numpy with the kernel's constants.
Checks, for every grid of eigen_init and ragged sizes, that
  * the ticket range covers every (strip, tile-row group) exactly once,
  * the reader sums exactly the partial entries the writer produced (no stale / missing entry),
  * the sum over all ranks is p = A u restricted to the strict upper triangle of A(0:L, 0:L) plus nothing else --
    the per-column exchange then adds the diagonal term and the panel corrections (src/eigen_trd_t2.F:970-1660)."""
import numpy as np
import pytest

TR, TC, SW, VS = 128, 64, 4, 8


def cyc_count(G, P, r):
    return (G - r + P - 1) // P if G > r else 0


class Geo:
    def __init__(self, L, px, py, x, y, G=296):
        self.L, self.px, self.py, self.x, self.y = L, px, py, x, y
        self.nclL = cyc_count(L, py, y)
        ntl = (cyc_count(L, px, x) // TR + 1) * (self.nclL // TC + 1) // 2
        per = ntl // (6 * G)
        self.sw = 4 if per >= 4 else 2 if per >= 2 else 1
        self.sv = 4 if per >= 16 else 2 if per >= 8 else 1

    def ntile_rows(self, t):
        clast = min((t + 1) * TC, self.nclL) - 1
        if clast < t * TC:
            return 0
        cmax_g = clast * self.py + self.y
        if cmax_g <= self.x:
            return 0
        return (cmax_g - self.x - 1) // (TR * self.px) + 1

    def nstrips(self):
        return (self.nclL + self.sw * TC - 1) // (self.sw * TC)

    def strip_rows(self, sc):
        tlast = min((sc + 1) * self.sw, (self.nclL + TC - 1) // TC) - 1
        return self.ntile_rows(tlast)

    def strip_groups(self, sc):
        return (self.strip_rows(sc) + self.sv - 1) // self.sv

    def fold(self, item, gx):
        nsc = self.nstrips()
        bx, by = item % gx, item // gx
        sc1, sc2 = nsc - 1 - bx, bx
        n1 = self.strip_groups(sc1)
        if by < n1:
            return sc1, by
        if sc2 == sc1:
            return None
        bg = by - n1
        return (sc2, bg) if bg < self.strip_groups(sc2) else None


def run_rank(A, u, L, px, py, x, y, force=None):
    g = Geo(L, px, py, x, y)
    if force:
        g.sw, g.sv = force
    a_loc = A[x::px, y::py]
    nrl_pad = (a_loc.shape[0] + TR - 1) // TR * TR + TR
    ncl_pad = (a_loc.shape[1] + TC * SW - 1) // (TC * SW) * TC * SW + TC * SW
    al = np.zeros((nrl_pad, ncl_pad))
    al[:a_loc.shape[0], :a_loc.shape[1]] = a_loc
    nsc = g.nstrips()
    gx = (nsc + 1) // 2
    gy = 0
    for bx in range(gx):
        s1, s2 = nsc - 1 - bx, bx
        gy = max(gy, g.strip_groups(s1) + (g.strip_groups(s2) if s2 != s1 else 0))
    Prow = np.full((max(nsc, 1), nrl_pad), np.nan)      # NaN = never written
    Pcol = np.full((nrl_pad // TR + 1, ncl_pad), np.nan)
    seen = set()
    for item in range(gx * gy):
        f = g.fold(item, gx)
        if f is None:
            continue
        sc, bg = f
        assert (sc, bg) not in seen
        seen.add((sc, bg))
        scol = np.zeros(g.sw * TC)
        for br in range(bg * g.sv, min(bg * g.sv + g.sv, g.strip_rows(sc))):
            r0 = br * TR
            acc_row = np.zeros(TR)
            for st in range(g.sw):
                t = sc * g.sw + st
                c0 = t * TC
                if c0 >= g.nclL:
                    break
                if br >= g.ntile_rows(t):
                    continue
                tile = al[r0:r0 + TR, c0:c0 + TC].copy()
                gr = (np.arange(r0, r0 + TR) * px + x)[:, None]
                gc = (np.arange(c0, c0 + TC) * py + y)[None, :]
                tile[~((gc < L) & (gr < gc))] = 0.0
                ux = np.where(gr[:, 0] < L, u[np.minimum(gr[:, 0], len(u) - 1)], 0.0)
                uy = np.where(gc[0] < L, u[np.minimum(gc[0], len(u) - 1)], 0.0)
                acc_row += tile @ uy
                scol[st * TC:(st + 1) * TC] += ux @ tile
            Prow[sc, r0:r0 + TR] = acc_row
        for st in range(g.sw):
            c0 = (sc * g.sw + st) * TC
            Pcol[bg, c0:c0 + TC] = scol[st * TC:(st + 1) * TC]
    for sc in range(nsc):
        for bg in range(g.strip_groups(sc)):
            assert (sc, bg) in seen, (sc, bg)
    # reader (p phase): rows this rank holds get the row partials of the strips that reach their tile row, columns
    # the column partials of their tile's groups
    s_nbr = [g.strip_rows(s) for s in range(nsc)]
    p = np.zeros(L)
    for gg in range(L):
        acc = 0.0
        if gg % px == x:
            jl = gg // px
            brr = jl // TR
            s_lo = next((s for s in range(nsc) if s_nbr[s] > brr), nsc)
            assert all(s_nbr[s] > brr for s in range(s_lo, nsc))          # monotone: what the binary search relies on
            for s in range(s_lo, nsc):
                assert not np.isnan(Prow[s, jl])
                acc += Prow[s, jl]
        if gg % py == y:
            il = gg // py
            nb = (g.ntile_rows(il // TC) + g.sv - 1) // g.sv
            for b in range(nb):
                assert not np.isnan(Pcol[b, il])
                acc += Pcol[b, il]
        p[gg] = acc
    return p


@pytest.mark.parametrize("px,py", [(1, 1), (1, 2), (2, 2), (2, 4), (3, 2)])
@pytest.mark.parametrize("L,force", [(5, None), (130, None), (700, None), (1111, None), (900, (2, 2)), (1500, (4, 4)), (1300, (4, 2))])
def test_partial_sums_reproduce_strict_upper_symv(px, py, L, force):
    n = L + 3
    rng = np.random.default_rng(L * 10 + px + py)
    A = rng.standard_normal((n, n))
    u = rng.standard_normal(n)
    p = np.zeros(L)
    for x in range(px):
        for y in range(py):
            p += run_rank(A, u, L, px, py, x, y, force)
    S = np.triu(A[:L, :L], 1)
    want = S @ u[:L] + S.T @ u[:L]
    assert np.allclose(p, want, rtol=1e-11, atol=1e-11)
