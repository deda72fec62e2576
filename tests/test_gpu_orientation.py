"""Non-aligned neighbours (OrientationMap): a periodic Brick whose elements are
given independently rotated / reflected logical frames must produce exactly the
same physics.  The GPU runs the rotated mesh (with neighbour directions and
face permutations), the oracle runs the aligned mesh; results are compared after
mapping back.  Reference: orient_variables_on_slice, Domain/Structure/
OrientationMapHelpers.cpp:25-120; call site ComputeTimeDerivative.hpp:712-721."""
import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, lib
from tests import rotation

pytestmark = pytest.mark.gpu


_signed_perms, _rotate_problem = rotation.signed_perms, rotation.rotate_problem


@pytest.mark.parametrize("system", ["gh", "sw"])
def test_rotated_blocks_match_aligned_mesh(system):
    N = 4
    rng = np.random.default_rng(99)
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    x, nbr = brick.coords(), brick.neighbors()
    J = brick.inverse_jacobian() + 0.1 * rng.uniform(-1, 1, (brick.n_elements, 9, N ** 3))
    if system == "gh":
        sysid, C = lib.SYSTEM_GH, 50
        u = analytic.gauge_wave(x, 0.1) + 1e-2 * rng.uniform(-1, 1, (brick.n_elements, 50, N ** 3))
        stat = rng.uniform(-1, 1, (brick.n_elements, 3, N ** 3))
        blocks = [slice(0, 10), slice(10, 20), slice(20, 50)]
    else:
        sysid, C = lib.SYSTEM_SCALAR_WAVE, 5
        u = analytic.plane_wave(x, 0.3) + 0.1 * rng.uniform(-1, 1, (brick.n_elements, 5, N ** 3))
        stat = rng.uniform(0, 1, (brick.n_elements, 1, N ** 3))
        blocks = [slice(0, 1), slice(1, 2), slice(2, 5)]
    all48 = _signed_perms()
    frames = [all48[k] for k in rng.choice(48, brick.n_elements, replace=False)]
    frames[0] = ((0, 1, 2), (1, 1, 1))  # keep one element aligned
    u_r, J_r, s_r, nbr_r, nd_r, perm_r, pm = _rotate_problem(N, u, J, stat, nbr, frames)
    assert (perm_r != 0).any() and (nd_r != (np.arange(6) ^ 1)[None, :]).any()
    ctx = lib.Context(sysid, N, brick.n_elements)
    ctx.set_geometry(J_r, None, nbr_r)
    with pytest.raises(lib.DgrhsError, match="not that of aligned blocks"):
        ctx.set_static_fields(s_r)
        ctx.set_state(u_r)
        ctx.compute_time_derivative(0.0)
    ctx.set_neighbor_orientations(nd_r, perm_r)
    ctx.compute_time_derivative(0.0)
    got_r = ctx.get_time_derivative()
    got = np.empty_like(got_r)
    for e in range(brick.n_elements):
        got[e] = got_r[e][:, pm[e]]
    ref = orc.dg_rhs(0 if system == "sw" else 1, N, u, J, stat, nbr)
    err = max(np.max(np.abs(got[:, b] - ref[:, b])) / np.max(np.abs(ref[:, b])) for b in blocks)
    assert err < 1e-12
    # a short evolution on the rotated mesh
    dt = 2e-4
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 2, 0.0, dt)
    ctx.take_steps(2)
    ev = orc.Evolution(lambda v, t: orc.dg_rhs(0 if system == "sw" else 1, N, v, J, stat, nbr),
                       u, 0.0, dt, "AB2")
    ev.step()
    ev.step()
    st_r = ctx.get_state()
    st = np.empty_like(st_r)
    for e in range(brick.n_elements):
        st[e] = st_r[e][:, pm[e]]
    err = max(np.max(np.abs(st[:, b] - ev.u[:, b])) / np.max(np.abs(ev.u[:, b])) for b in blocks)
    assert err < 1e-12
    ctx.close()


def test_orientation_validation():
    N = 3
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(brick.inverse_jacobian(), None, brick.neighbors())
    nd = np.tile(np.arange(6) ^ 1, (brick.n_elements, 1))
    pm = np.zeros_like(nd)
    ctx.set_neighbor_orientations(nd, pm)  # the aligned default, explicitly
    pm[0, 0] = 1  # a swap on one side only is not an inverse pair
    with pytest.raises(lib.DgrhsError, match="not inverse"):
        ctx.set_neighbor_orientations(nd, pm)
    ctx.close()
