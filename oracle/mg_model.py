"""NumPy model of the geometric multigrid V-cycle that replaces the reference's
poisson_multigrid (TEST INFRASTRUCTURE ONLY -- same import rules as the rest of oracle/).

The reference routine (src/poisson_multigrid.f90:10-189) is undefined behaviour as called
(DESIGN.md section 6), so there is no reference result to restate.  This module is a CPU model of
the design that csrc/multigrid.cu implements -- same level hierarchy, same 1-D transfer tables
-- used (a) to check the device transfer operators component by component and (b) to study
convergence on the shipped grid shapes.  MG *parity* is pinned elsewhere: against the SOR
oracle (src/poisson.f90 restatement) on the same operator, modulo the additive constant.

Operator and neighbour rule: src/poisson.f90:42-51,57-92 (periodic wrap / mirror).
"""
import numpy as np

WRAP, MIRROR = 0, 1
MIN_N = 5          # an axis with fewer than this many points is not coarsened further


def coarse_extent(n, mode):
    """points of the next coarser level along one axis (0 = do not coarsen)"""
    if mode == MIRROR:
        nc = (n + 1) // 2 if n % 2 else n // 2 + 1
    else:
        nc = n // 2 if n % 2 == 0 else (n + 1) // 2
    if n < MIN_N or nc < 3:
        return 0
    return nc


def axis_tables(n, d, mode, nc):
    """1-D transfer tables fine(n, d) <-> coarse(nc, D).
    returns D, c0[n], w[n] (prolongation: f(i) = (1-w) C[c0] + w C[c1], c1 = next(c0)),
    ridx[nc,4], rw[nc,4] (restriction: C[c] = sum_t rw[c,t] f(ridx[c,t]))."""
    if mode == MIRROR:
        L = (n - 1) * d
        D = L / (nc - 1)
    else:
        L = n * d
        D = L / nc
    c0 = np.zeros(n, dtype=np.int32)
    w = np.zeros(n)
    for i in range(n):
        # exact integer arithmetic for the nested cases, so weights are exactly 0 / 0.5
        if mode == MIRROR:
            num, den = i * (nc - 1), (n - 1)
        else:
            num, den = i * nc, n
        c = num // den
        frac = (num - c * den) / den
        if mode == MIRROR and c >= nc - 1:
            c, frac = nc - 2, 1.0
        c0[i], w[i] = c, frac
    # transpose on the even / periodic extension, rows normalised to sum 1
    rows = [dict() for _ in range(nc)]
    for i in range(n):
        c = int(c0[i])
        c1 = c + 1
        if mode == WRAP:
            c1 %= nc
        for cc, ww in ((c, 1.0 - w[i]), (c1, w[i])):
            if ww != 0.0:
                rows[cc][i] = rows[cc].get(i, 0.0) + ww
    if mode == MIRROR:
        # wall coarse nodes also collect the mirror images of the off-wall fine nodes
        for cc, wall in ((0, 0), (nc - 1, n - 1)):
            for i in list(rows[cc]):
                if i != wall:
                    rows[cc][i] *= 2.0
    ridx = np.zeros((nc, 4), dtype=np.int32)
    rw = np.zeros((nc, 4))
    for c in range(nc):
        items = sorted(rows[c].items())
        assert 1 <= len(items) <= 4, (n, nc, mode, c, items)
        s = sum(v for _, v in items)
        for t, (i, v) in enumerate(items):
            ridx[c, t], rw[c, t] = i, v / s
        for t in range(len(items), 4):
            ridx[c, t] = items[0][0]
    return D, c0, w, ridx, rw


class Level:
    def __init__(self, n, d, modes):
        self.n, self.d, self.modes = tuple(n), tuple(d), tuple(modes)
        self.o = [1.0 / (x * x) for x in d]                      # src/poisson.f90:42-47
        self.A = -(2.0 * self.o[0] + 2.0 * self.o[1] + 2.0 * self.o[2])   # :48-51
        self.tab = None     # tables to the next coarser level

    def nbr(self, ax):
        n, mode = self.n[ax], self.modes[ax]
        i = np.arange(n)
        m, p = i - 1, i + 1
        if mode == WRAP:
            m[0], p[-1] = n - 1, 0
        else:
            m[0], p[-1] = 1, n - 2
        return m, p

    def apply(self, p):
        out = self.A * p
        for ax in range(3):
            m, q = self.nbr(ax)
            out += self.o[ax] * (np.take(p, m, axis=ax) + np.take(p, q, axis=ax))
        return out

    def classes(self):
        """(colour, seam parity) classes in sweep order; each is an independent set"""
        n = self.n
        I, J, K = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
        col = (I + J + K) & 1
        seam = np.zeros_like(col)
        for ax, idx in enumerate((I, J, K)):
            if self.modes[ax] == WRAP and n[ax] % 2:
                seam += (idx == n[ax] - 1)
        seam &= 1
        out = []
        for sp in (0, 1):
            for c in (0, 1):
                mask = (col == c) & (seam == sp)
                if mask.any():
                    out.append(mask)
        return out

    def smooth(self, p, rhs, sweeps, omega=1.0):
        cls = self.classes()
        for _ in range(sweeps):
            for mask in cls:
                s = np.zeros_like(p)
                for ax in range(3):
                    m, q = self.nbr(ax)
                    s += self.o[ax] * (np.take(p, m, axis=ax) + np.take(p, q, axis=ax))
                pn = (rhs - s) / self.A
                p[mask] = ((1.0 - omega) * p + omega * pn)[mask]
        return p


def build_hierarchy(n, d, modes, max_levels=32):
    levels = [Level(n, d, modes)]
    while len(levels) < max_levels:
        f = levels[-1]
        nc = [coarse_extent(f.n[a], f.modes[a]) for a in range(3)]
        if not any(nc):
            break
        tabs, cn, cd = [], [], []
        for a in range(3):
            if nc[a]:
                D, c0, w, ridx, rw = axis_tables(f.n[a], f.d[a], f.modes[a], nc[a])
                cn.append(nc[a])
            else:   # identity along this axis
                D = f.d[a]
                c0 = np.arange(f.n[a], dtype=np.int32)
                w = np.zeros(f.n[a])
                ridx = np.repeat(np.arange(f.n[a], dtype=np.int32)[:, None], 4, axis=1)
                rw = np.zeros((f.n[a], 4))
                rw[:, 0] = 1.0
                cn.append(f.n[a])
            cd.append(D)
            tabs.append((c0, w, ridx, rw))
        f.tab = tabs
        levels.append(Level(cn, cd, modes))
    return levels


def restrict(f, c, r):
    out = r
    for ax in range(3):
        _, _, ridx, rw = f.tab[ax]
        acc = 0.0
        for t in range(4):
            shp = [1, 1, 1]
            shp[ax] = -1
            acc = acc + np.take(out, ridx[:, t], axis=ax) * rw[:, t].reshape(shp)
        out = acc
    return out


def prolong(f, c, e):
    out = e
    for ax in range(3):
        c0, w, _, _ = f.tab[ax]
        c1 = c0 + 1
        if f.modes[ax] == WRAP:
            c1 = c1 % c.n[ax]
        else:
            c1 = np.minimum(c1, c.n[ax] - 1)
        shp = [1, 1, 1]
        shp[ax] = -1
        out = np.take(out, c0, axis=ax) * (1.0 - w).reshape(shp) + \
            np.take(out, c1, axis=ax) * w.reshape(shp)
    return out


def wall_weights(lv):
    """left null vector of the operator: 1/2 per mirrored wall the point lies on"""
    w = np.ones(lv.n)
    for ax in range(3):
        if lv.modes[ax] == MIRROR:
            sl = [slice(None)] * 3
            for e in (0, lv.n[ax] - 1):
                sl[ax] = e
                w[tuple(sl)] *= 0.5
    return w


def vcycle(levels, l, p, rhs, npre, npost, ncoarse=40):
    lv = levels[l]
    if l == len(levels) - 1:
        w = wall_weights(lv)
        rhs = rhs - np.sum(rhs * w) / np.sum(w)     # compatibility of the singular system
        return lv.smooth(p, rhs, ncoarse)
    lv.smooth(p, rhs, npre)
    r = rhs - lv.apply(p)
    rc = restrict(lv, levels[l + 1], r)
    ec = vcycle(levels, l + 1, np.zeros(levels[l + 1].n), rc, npre, npost, ncoarse)
    p += prolong(lv, levels[l + 1], ec)
    lv.smooth(p, rhs, npost)
    return p


def solve(p, rhs, d, modes, npre=5, npost=4, tol=1e-8, max_cycles=50, verbose=False):
    levels = build_hierarchy(p.shape, d, modes)
    hist = []
    for cyc in range(max_cycles):
        r = rhs - levels[0].apply(p)
        dmax = np.max(np.abs(r)) / abs(levels[0].A)
        hist.append(dmax)
        if verbose:
            print("cycle %d dmax %.3e" % (cyc, dmax))
        if dmax < tol:
            break
        if cyc >= 2 and dmax > 0.9 * hist[-2]:
            break
        vcycle(levels, 0, p, rhs, npre, npost)
    return p, hist, levels
