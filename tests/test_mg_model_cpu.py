"""CPU checks of the multigrid design model (oracle/mg_model.py): the 1-D transfer tables that
csrc/multigrid.cu builds the same way, and the convergence of the V(5,4) cycle on nested and
non-nested hierarchies (odd periodic extents, even mirrored extents, mixed boundary rules)."""
import numpy as np
import pytest

from oracle import mg_model as M


@pytest.mark.parametrize("n,mode", [(33, 1), (32, 1), (32, 0), (33, 0), (241, 0), (81, 0), (6, 1)])
def test_transfer_tables(n, mode):
    nc = M.coarse_extent(n, mode)
    assert nc >= 3
    D, c0, w, ridx, rw = M.axis_tables(n, 0.1, mode, nc)
    assert np.all((w >= 0) & (w <= 1)) and np.all((c0 >= 0) & (c0 < nc))
    assert np.allclose(rw.sum(axis=1), 1.0, atol=1e-15)
    assert np.all(rw >= 0)
    # constants are reproduced by interpolation and by restriction
    nested = (mode == 1 and n % 2 == 1) or (mode == 0 and n % 2 == 0)
    if nested:
        assert abs(D - 0.2) < 1e-15
        assert set(np.unique(w)) <= {0.0, 0.5, 1.0}
        inner = rw[1:-1] if mode == 1 else rw
        assert np.allclose(np.sort(inner, axis=1)[:, 1:], [0.25, 0.25, 0.5])
    # a linear function is interpolated exactly on a mirrored axis
    if mode == 1:
        xc = D * np.arange(nc)
        xf = 0.1 * np.arange(n)
        c1 = np.minimum(c0 + 1, nc - 1)
        assert np.allclose((1 - w) * xc[c0] + w * xc[c1], xf, atol=1e-13)


@pytest.mark.parametrize("shape,modes", [((17, 17, 17), (1, 1, 1)), ((16, 18, 14), (1, 1, 1)),
                                         ((16, 16, 16), (0, 0, 0)), ((17, 15, 13), (0, 0, 0)),
                                         ((21, 18, 11), (0, 1, 0))])
def test_vcycle_converges(shape, modes):
    rng = np.random.default_rng(0)
    d = (0.1, 0.11, 0.09)
    lv = M.Level(shape, d, modes)
    pt = rng.standard_normal(shape)
    rhs = lv.apply(pt)
    tol = 1e-10 * np.max(np.abs(rhs)) / abs(lv.A)
    p, hist, levels = M.solve(np.zeros(shape), rhs, d, modes, 5, 4, tol, 30)
    assert hist[-1] < tol and len(hist) - 1 <= 12, hist
    rates = [hist[i + 1] / hist[i] for i in range(1, len(hist) - 1)]
    assert max(rates) < 0.15, rates
    assert np.max(np.abs((p - p.mean()) - (pt - pt.mean()))) < 1e-6
