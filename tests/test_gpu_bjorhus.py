"""ConstraintPreservingBjorhus (Types ConstraintPreserving and
ConstraintPreservingPhysical) through the C-ABI:
gh_bjorhus_kernel against the oracle (oracle/bjorhus.py, pinned to the
reference's Bjorhus.py), on a Brick and on the outer boundary of the Kerr-Schild
shell.  Reference: GeneralizedHarmonic/BoundaryConditions/Bjorhus.cpp,
BjorhusImpl.cpp; applied as a TimeDerivative-type condition by
BoundaryConditionsImpl.hpp:566-670."""
import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, evolution, lib

pytestmark = pytest.mark.gpu

TOL = 1e-12
GH_BLOCKS = [slice(0, 10), slice(10, 20), slice(20, 50)]


def _relerr(a, b, blocks):
    return max(np.max(np.abs(a[:, s] - b[:, s])) / np.max(np.abs(b[:, s])) for s in blocks)


@pytest.mark.parametrize("N,kind", [(3, lib.BJORHUS), (6, lib.BJORHUS), (4, lib.BJORHUS_PHYSICAL),
                                    (7, lib.BJORHUS_PHYSICAL)])
def test_bjorhus_on_brick_faces_matches_oracle(N, kind):
    """Gauge wave on a Brick away from the origin, periodic in y and z, Bjorhus on
    both x faces, harmonic gauge, perturbed state and Jacobian."""
    rng = np.random.default_rng(N)
    brick = domain.Brick([3.0, -0.5, -0.5], [4.0, 0.5, 0.5], [1, 1, 1], N,
                         periodic=(False, True, True))
    x, nbr = brick.coords(), brick.neighbors().copy()
    assert (nbr == -1).sum() == 8
    nbr[nbr == -1] = kind
    J = brick.inverse_jacobian() + 0.05 * rng.uniform(-1, 1, (brick.n_elements, 9, N ** 3))
    u = analytic.gauge_wave(x, 0.1) + 1e-2 * rng.uniform(-1, 1, (brick.n_elements, 50, N ** 3))
    stat = rng.uniform(-1, 1, (brick.n_elements, 3, N ** 3))
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, x, nbr)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(1, N, u, J, stat, nbr, coords=x)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    # the boundary condition matters
    plain = orc.dg_rhs(1, N, u, J, stat, np.where(nbr == kind, -1, nbr), coords=x)
    other = lib.BJORHUS if kind == lib.BJORHUS_PHYSICAL else lib.BJORHUS_PHYSICAL
    assert _relerr(orc.dg_rhs(1, N, u, J, stat, np.where(nbr == kind, other, nbr), coords=x),
                   ref, GH_BLOCKS) > 1e-3      # the two types differ
    assert _relerr(plain, ref, GH_BLOCKS) > 1e-3
    # and an AB2 evolution
    dt = 1e-4
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 2, 0.0, dt)
    ctx.take_steps(2)
    ev = orc.Evolution(lambda v, t: orc.dg_rhs(1, N, v, J, stat, nbr, coords=x), u, 0.0, dt, "AB2")
    ev.step()
    ev.step()
    assert _relerr(ctx.get_state(), ev.u, GH_BLOCKS) < TOL
    ctx.close()


@pytest.mark.parametrize("kind", ["ConstraintPreserving", "ConstraintPreservingPhysical"])
def test_bjorhus_on_kerr_schild_shell(kind):
    """Kerr-Schild shell with DemandOutgoingCharSpeeds on the excision sphere and
    ConstraintPreservingBjorhus on the outer sphere (the boundary conditions of
    the single-black-hole executables), AnalyticChristoffel gauge (gauge fields)."""
    N = 5
    problem = evolution.gh_kerr_schild_shell_problem(
        (0, 0), N, inner_radius=1.9, outer_radius=6.0,
        inner_boundary="DemandOutgoingCharSpeeds", outer_boundary=kind)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ctx, part = ev.ctx, ev.part
    code = lib.BJORHUS_PHYSICAL if kind == "ConstraintPreservingPhysical" else lib.BJORHUS
    assert (part.local_neighbors == code).sum() == 6 and not part.external_faces
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    u = u0 + 1e-3 * np.random.default_rng(5).uniform(-1, 1, u0.shape)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    H = np.zeros((len(ids), 4, N ** 3))
    dH = np.zeros((len(ids), 16, N ** 3))
    for e in range(len(ids)):
        H[e], dH[e] = orc.analytic_christoffel_gauge(N, u0[e], J[e])
    sf = np.concatenate([stat, H, dH], axis=1)
    ref = orc.dg_rhs(1, N, u, J, sf, part.local_neighbors, gauge_params=orc.GAUGE_GIVEN,
                     coords=x, nbr_dir=part.local_neighbor_direction,
                     face_perm=part.local_face_permutation)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    # on the exact solution the condition only sees truncation-level constraint
    # violations: the static black hole stays put
    ctx.set_state(u0)
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, 1e-4)
    ev.take_steps(3)
    ctx.check_outgoing_char_speeds()
    assert np.max(np.abs(ctx.get_state() - u0)) < 1e-3
    ctx.close()


def test_bjorhus_with_damped_harmonic_gauge():
    """The boundary conditions of the binary-black-hole input files:
    ConstraintPreservingPhysical with the DampedHarmonic gauge (H_a and d_a H_b are
    evaluated at the face points by the same code as in the volume kernel)."""
    N = 5
    rng = np.random.default_rng(77)
    brick = domain.Brick([3.0, -0.5, -0.5], [4.0, 0.5, 0.5], [1, 1, 1], N,
                         periodic=(False, True, True))
    x, nbr = brick.coords(), brick.neighbors().copy()
    nbr[nbr == -1] = lib.BJORHUS_PHYSICAL
    J = brick.inverse_jacobian() + 0.05 * rng.uniform(-1, 1, (brick.n_elements, 9, N ** 3))
    u = analytic.gauge_wave(x, 0.1) + 1e-2 * rng.uniform(-1, 1, (brick.n_elements, 50, N ** 3))
    stat = rng.uniform(-1, 1, (brick.n_elements, 3, N ** 3))
    params = [3.0, 1.2, 1.5, 1.7, 2, 4, 6]
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, x, nbr)
    ctx.set_static_fields(stat)
    ctx.set_gauge(lib.GAUGE_DAMPED_HARMONIC, params)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    gp = np.array([2.0] + params)
    ref = orc.dg_rhs(1, N, u, J, stat, nbr, gauge_params=gp, coords=x)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    harmonic = orc.dg_rhs(1, N, u, J, stat, nbr, coords=x)
    assert _relerr(harmonic, ref, GH_BLOCKS) > 1e-9   # the gauge source does enter
    ctx.close()


def test_bjorhus_misuse():
    N = 3
    brick = domain.Brick([3.0, 0, 0], [4.0, 1, 1], [0, 0, 0], N, periodic=(False,) * 3)
    nbr = np.full((1, 6), lib.BJORHUS, dtype=np.int32)
    sw = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, 1)
    with pytest.raises(lib.DgrhsError, match="GeneralizedHarmonic boundary condition"):
        sw.set_geometry(brick.inverse_jacobian(), brick.coords(), nbr)
    sw.close()
    gh = lib.Context(lib.SYSTEM_GH, N, 1)
    with pytest.raises(lib.DgrhsError, match="needs inertial coordinates"):
        gh.set_geometry(brick.inverse_jacobian(), None, nbr)
    # faces that share edge points would have to be applied one after the other
    # (BoundaryConditionsImpl.hpp:277-278, 636-660): rejected, opposite faces are fine
    with pytest.raises(lib.DgrhsError, match="more than one dimension"):
        gh.set_geometry(brick.inverse_jacobian(), brick.coords(), nbr)
    opposite = np.array([[lib.BJORHUS, lib.BJORHUS, -1, -1, -1, -1]], dtype=np.int32)
    gh.set_geometry(brick.inverse_jacobian(), brick.coords(), opposite)
    gh.close()
