"""Dirichlet wall closure (O3D_CLOSURE_D11) -- a NEW feature with NO reference parity: the north
star and the reference's README (README.md:28,123-124) name a Dirichlet-x closure, the source has
none (src/initialization.f90:228-242 stops on anything but PERIODIC / FREE_SLIP).  Ghost rule
f(1-g) = 2 f(1) - f(1+g) (odd reflection about the stored wall value; the reference's odd closure
der?i_11 is its f_wall = 0 case).  Checked the only way possible, analytically:

  * it degenerates to der?i_11 BIT FOR BIT when the wall planes hold zero, and equals
    der?i_11(f - f_wall) + 0 for a constant wall value (linearity of the stencil);
  * manufactured solutions: exact for polynomials the interior stencil differentiates exactly and
    whose odd extension about the wall value is smooth; convergence order on a smooth function
    with f'' = 0 at the walls (where the reflected extension is C^3): 6th-order interior, the
    wall rows limited by the extension's smoothness -- measured and asserted."""
import numpy as np
import pytest

from conftest import rand_field

pytestmark = pytest.mark.gpu
D11, I11 = 4, 2


@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("order", [1, 2])
def test_dirichlet_with_zero_wall_is_the_odd_closure_bitwise(gpu, axis, order):
    from osinco3d_b200 import modules as M
    f = rand_field((23, 19, 17), 3)
    idx = [slice(None)] * 3
    idx[axis] = [0, -1]
    f[tuple(idx)] = 0.0
    a = M.der(axis, order, D11, f, 0.05)
    b = M.der(axis, order, I11, f, 0.05)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_dirichlet_constant_wall_value_shifts_out(gpu, axis):
    """f = c + g with g odd about the walls: d/dx sees g only (the stencil annihilates constants,
    and 2 c - (c + g) = c - g is the odd image of g shifted by c)"""
    from osinco3d_b200 import modules as M
    g = rand_field((21, 18, 16), 5)
    idx = [slice(None)] * 3
    idx[axis] = [0, -1]
    g[tuple(idx)] = 0.0
    c = 3.25
    for order in (1, 2):
        a = M.der(axis, order, D11, np.asfortranarray(g + c), 0.1)
        b = M.der(axis, order, I11, g, 0.1)
        assert np.max(np.abs(a - b)) <= 2e-12 * max(1.0, np.max(np.abs(b)))


def test_dirichlet_x_manufactured_solution_and_order(gpu):
    """u(x) = U0 + A x + sin(2 pi x / L) on [0, L]: u'' = 0 at both walls, so the reflection about
    the wall values is smooth; the first derivative converges at the interior scheme's order and
    the linear part is differentiated exactly"""
    from osinco3d_b200 import modules as M
    L, U0, A = 2.0, 0.7, -1.3
    # exactness: linear profile with non-zero wall values
    n = 17
    x = np.linspace(0.0, L, n)
    lin = np.asfortranarray((U0 + A * x)[:, None, None] * np.ones((n, 8, 8)))
    d = x[1] - x[0]
    assert np.max(np.abs(M.der(0, 1, D11, lin, d) - A)) < 1e-12
    assert np.max(np.abs(M.der(0, 2, D11, lin, d))) < 1e-10
    errs1, errs2 = [], []
    for n in (33, 65, 129):
        x = np.linspace(0.0, L, n)
        d = x[1] - x[0]
        u = U0 + A * x + np.sin(2 * np.pi * x / L)
        f = np.asfortranarray(u[:, None, None] * np.ones((n, 8, 8)))
        du = A + (2 * np.pi / L) * np.cos(2 * np.pi * x / L)
        d2u = -(2 * np.pi / L) ** 2 * np.sin(2 * np.pi * x / L)
        errs1.append(np.max(np.abs(M.der(0, 1, D11, f, d)[:, 0, 0] - du)))
        errs2.append(np.max(np.abs(M.der(0, 2, D11, f, d)[:, 0, 0] - d2u)))
    o1 = [np.log2(errs1[q] / errs1[q + 1]) for q in range(2)]
    o2 = [np.log2(errs2[q] / errs2[q + 1]) for q in range(2)]
    print("Dirichlet-x closure: first-derivative orders %s, second-derivative orders %s" % (o1, o2))
    assert min(o1) > 5.5 and min(o2) > 3.5


def test_dirichlet_rejects_nothing_else(gpu):
    from osinco3d_b200 import modules as M
    with pytest.raises(gpu.O3DError):
        M.der(0, 1, 5, rand_field((9, 9, 9), 1), 0.1)
