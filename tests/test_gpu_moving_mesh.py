"""Moving-mesh terms (`mesh_velocity`; VolumeTermsImpl.tpp:155-235, GH
TimeDerivative.cpp:237-300,372-378, normal_dot_mesh_velocity in dg_package_data): the CUDA
path with dgrhs_set_mesh_velocity against the oracle, 1e-12."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, evolution, lib
from tests.test_gpu_parity import (GH_BLOCKS, SW_BLOCKS, TOL, _curved_jacobian, _gh_problem,
                                   _relerr)

pytestmark = pytest.mark.gpu


def _periodic_velocity(x, L, amp):
    """A smooth velocity field that is periodic over the brick (continuous across every
    interface, like the velocity of a time-dependent map)."""
    k = 2 * np.pi / L
    v = np.empty((x.shape[0], 3, x.shape[2]))
    v[:, 0] = amp * (0.3 + np.sin(k * x[:, 1]) * np.cos(k * x[:, 2]))
    v[:, 1] = amp * (-0.2 + np.sin(k * x[:, 2] + 0.4) * np.cos(k * x[:, 0]))
    v[:, 2] = amp * (0.1 + np.sin(k * x[:, 0] + 1.1) * np.cos(k * x[:, 1]))
    return v


def test_package_data_operators_with_mesh_velocity(golden_dir):
    """dg_package_data with normal_dot_mesh_velocity: GH against the fixtures made by the
    reference's UpwindPenalty.py twin, ScalarWave against the oracle."""
    z = np.load(os.path.join(golden_dir, "upwind_penalty.npz"))
    m = np.load(os.path.join(golden_dir, "upwind_penalty_moving.npz"))
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    pk, speed = lib.gh_package_data(c(z["gh_u"][0].T), z["gh_gamma1"][0], z["gh_gamma2"][0],
                                    z["gh_lapse"][0], c(z["gh_shift"][0].T), c(z["gh_nlo"][0].T),
                                    c(z["gh_nup"][0].T), m["gh_ndotv"])
    np.testing.assert_allclose(pk.T, m["gh_packaged"], rtol=1e-13, atol=1e-14)
    assert speed == np.max(m["gh_packaged"][:, 130:])
    ref = orc.sw_package_data_moving(c(z["sw_u"][0].T), z["sw_gamma2"][0],
                                     c(z["sw_normal"][0].T), m["sw_ndotv"])
    pk, speed = lib.sw_package_data(c(z["sw_u"][0].T), z["sw_gamma2"][0], c(z["sw_normal"][0].T),
                                    m["sw_ndotv"])
    np.testing.assert_allclose(pk, ref, rtol=1e-13, atol=1e-14)
    assert speed == np.max(ref[13:])


@pytest.mark.parametrize("N,refine", [(3, 1), (5, 1), (8, 1), (12, 1)])
def test_scalar_wave_rhs_moving_mesh(N, refine):
    rng = np.random.default_rng(500 + N)
    L = 2 * np.pi
    brick = domain.Brick([0, 0, 0], [L] * 3, [refine] * 3, N)
    x = brick.coords()
    u = analytic.plane_wave(x, 0.3) + 0.1 * rng.uniform(-1, 1, (brick.n_elements, 5, brick.n))
    J = _curved_jacobian(rng, brick)
    nb = brick.neighbors()
    stat = rng.uniform(0, 1, (brick.n_elements, 1, brick.n))
    v = _periodic_velocity(x, L, 0.9)   # |n.v| crosses 1: every upwind weight switches somewhere
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    ctx.set_mesh_velocity(v)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(0, N, u, J, stat, nb, mesh_velocity=v)
    static = orc.dg_rhs(0, N, u, J, stat, nb)
    assert _relerr(got, ref, SW_BLOCKS) < TOL
    assert _relerr(static, ref, SW_BLOCKS) > 1e-3   # the terms matter
    # the stepper on a moving mesh (no fused update), then back to a static mesh
    dt = 1e-3
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
    ctx.take_steps(2)
    ev = orc.Evolution(lambda w, t: orc.dg_rhs(0, N, w, J, stat, nb, mesh_velocity=v), u, 0.0,
                       dt, "AB3")
    ev.step()
    ev.step()
    assert _relerr(ctx.get_state(), ev.u, SW_BLOCKS) < TOL
    ctx.set_mesh_velocity(None)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    assert _relerr(ctx.get_time_derivative(), static, SW_BLOCKS) < TOL
    ctx.close()


@pytest.mark.parametrize("N,refine,gauge", [(4, 1, "harmonic"), (6, 1, "fields"), (8, 1, "harmonic"),
                                            (10, 1, "fields"), (12, 1, "fields")])
def test_gh_rhs_moving_mesh(N, refine, gauge):
    rng = np.random.default_rng(600 + N)
    brick, x, u, J, stat = _gh_problem(rng, N, refine)
    nb = brick.neighbors()
    v = _periodic_velocity(x, 1.0, 0.7)
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    if gauge == "fields":
        H = rng.uniform(-1, 1, (brick.n_elements, 4, brick.n))
        dH = rng.uniform(-1, 1, (brick.n_elements, 16, brick.n))
        ctx.set_gauge(lib.GAUGE_FIELDS)
        ctx.set_gauge_fields(H, dH)
        ostat, gp = np.concatenate([stat, H, dH], axis=1), orc.GAUGE_GIVEN
    else:
        ostat, gp = stat, orc.GAUGE_HARMONIC
    ctx.set_state(u)
    ctx.set_mesh_velocity(v)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(1, N, u, J, ostat, nb, gauge_params=gp, mesh_velocity=v)
    static = orc.dg_rhs(1, N, u, J, ostat, nb, gauge_params=gp)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    assert _relerr(static, ref, GH_BLOCKS) > 1e-3
    if N <= 8:
        dt = 2e-4
        ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
        ctx.take_steps(2)
        ev = orc.Evolution(lambda w, t: orc.dg_rhs(1, N, w, J, ostat, nb, gauge_params=gp,
                                                   mesh_velocity=v), u, 0.0, dt, "AB3")
        ev.step()
        ev.step()
        assert _relerr(ctx.get_state(), ev.u, GH_BLOCKS) < TOL
    ctx.close()


def test_gh_moving_shell_with_ghost_boundaries():
    """Kerr-Schild shell (non-aligned wedges, DirichletAnalytic ghosts on both spheres) with
    the velocity of a rotating, expanding grid: v = Omega x r + a r."""
    from tests.test_gpu_shell import _gauge_fields
    N = 5
    problem = evolution.gh_kerr_schild_shell_problem((0, 0), N)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ctx, part = ev.ctx, ev.part
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    rng = np.random.default_rng(N)
    u = u0 + 1e-3 * rng.uniform(-1, 1, u0.shape)
    omega = np.array([0.02, -0.05, 0.3])
    v = np.cross(omega[None, :, None], x, axis=1) + 0.05 * x
    ctx.set_state(u)
    ctx.set_mesh_velocity(v)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    H, dH = _gauge_fields(N, x, J, u0)
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    ref = orc.dg_rhs(1, N, u, J, np.concatenate([stat, H, dH], axis=1), part.local_neighbors,
                     gauge_params=orc.GAUGE_GIVEN, ext_u=ext,
                     nbr_dir=part.local_neighbor_direction,
                     face_perm=part.local_face_permutation, mesh_velocity=v)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    ctx.close()


def test_moving_mesh_rejects_unsupported_faces():
    N = 4
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [0, 0, 0], N, periodic=(True, True, False))
    nb = brick.neighbors().copy()
    nb[nb == -1] = lib.BJORHUS
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(brick.inverse_jacobian(), brick.coords(), nb)
    with pytest.raises(lib.DgrhsError, match="moving mesh"):
        ctx.set_mesh_velocity(np.zeros((brick.n_elements, 3, brick.n)))
    ctx.close()
