"""Structural cross-checks between the 18 stencil routines of the oracle (src/derivation.f90),
independent of any stored vector: periodic routines commute with cyclic shifts; the free-slip
routines are the periodic routine applied to the even / odd mirror extension of the line; the
schemes have their nominal order of accuracy; the divergence guard of correct_velocity fires
exactly on NaN or a value above 1000 (src/integration.f90:309-325)."""
import numpy as np
import pytest

from conftest import rand_field

SHAPE = (11, 9, 13)
D = 0.043


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("order", [1, 2])
def test_periodic_routines_commute_with_cyclic_shifts(O, axis, order):
    f = rand_field(SHAPE, 5)
    ref = O.der(axis, order, 0, f, D)
    for shift in (1, 3, SHAPE[axis] - 2):
        got = O.der(axis, order, 0, np.asfortranarray(np.roll(f, shift, axis)), D)
        assert np.array_equal(got, np.roll(ref, shift, axis)), (axis, order, shift)


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("closure", [1, 2])
def test_free_slip_routines_are_the_periodic_routine_on_the_mirror_extension(O, axis, order, closure):
    """f(1-k) = +-f(1+k), f(n+k) = +-f(n-k): a line of n points extends to a periodic line of
    2n - 2 points; x - (-y) == x + y bitwise, so even the boundary planes agree bit for bit"""
    f = rand_field(SHAPE, 6)
    n = SHAPE[axis]
    sign = 1.0 if closure == 1 else -1.0
    idx = [slice(None)] * 3
    idx[axis] = slice(n - 2, 0, -1)
    ext = np.asfortranarray(np.concatenate([f, sign * f[tuple(idx)]], axis=axis))
    per = O.der(axis, order, 0, ext, D)
    idx[axis] = slice(0, n)
    assert np.array_equal(O.der(axis, order, closure, f, D), per[tuple(idx)])


@pytest.mark.parametrize("order,slope", [(1, 6.0), (2, 4.0)])
def test_order_of_accuracy(O, order, slope):
    errs = []
    for n in (16, 32, 64):
        d = 2 * np.pi / n                      # n points per period: exactly periodic samples
        x = d * np.arange(n)
        f = np.asfortranarray(np.sin(x)[:, None, None] * np.ones((1, 7, 7)))
        exact = (np.cos(x) if order == 1 else -np.sin(x))[:, None, None] * np.ones((1, 7, 7))
        errs.append(np.max(np.abs(O.der(0, order, 0, f, d) - exact)))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert all(abs(r - slope) < 0.3 for r in rates), (errs, rates)


def test_divergence_guard_thresholds(O):
    g = O.grid(*SHAPE, D, D, D, (1, 1, 1))
    z = np.asfortranarray(np.zeros(SHAPE))
    u = [rand_field(SHAPE, s, 0.1) for s in (1, 2, 3)]
    assert O.correct_velocity(g, *u, z, 1e-3)[3] == 0
    for comp in range(3):
        for val, bad in ((1000.0, 0), (1000.0000001, 1), (-5000.0, 0), (np.nan, 1)):
            v = [a.copy(order="F") for a in u]
            v[comp][4, 5, 6] = val
            assert (O.correct_velocity(g, *v, z, 1e-3)[3] != 0) == bool(bad), (comp, val)
