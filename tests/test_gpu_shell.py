"""BASELINE.json configs[2] on its real geometry: GeneralizedHarmonic Kerr-Schild
on the spherical shell of tests/InputFiles/GeneralizedHarmonic/KerrSchild.yaml
(six equiangular Wedge<3> blocks per layer, non-aligned neighbours between the
blocks, full 3x3 inverse Jacobian per point, excision boundary).  The GPU path
goes through the C-ABI; the checker is the oracle with orient_variables_on_slice
and the DirichletAnalytic ghost state."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, evolution, lib

pytestmark = pytest.mark.gpu

TOL = 1e-12
GH_BLOCKS = [slice(0, 10), slice(10, 20), slice(20, 50)]
SW_BLOCKS = [slice(0, 1), slice(1, 2), slice(2, 5)]


def _relerr(a, b, blocks):
    return max(np.max(np.abs(a[:, s] - b[:, s])) / np.max(np.abs(b[:, s])) for s in blocks)


def _gauge_fields(N, x, J, u0):
    H = np.zeros((len(x), 4, N ** 3))
    dH = np.zeros((len(x), 16, N ** 3))
    for e in range(len(x)):
        ua = orc.gh_vars_from_metric(*orc.kerr_schild_metric(x[e]))
        np.testing.assert_allclose(ua, u0[e], rtol=1e-13, atol=1e-14)
        H[e], dH[e] = orc.analytic_christoffel_gauge(N, ua, J[e])
    return H, dH


@pytest.mark.parametrize("N,refinement,layers", [(5, (0, 0), ()), (4, (1, 1), (2.1,)),
                                                 (8, (1, 0), ()), (10, (0, 1), ()),
                                                 (12, (0, 0), ()), (12, (1, 0), ())])
def test_gh_rhs_on_shell_matches_oracle(N, refinement, layers):
    problem = evolution.gh_kerr_schild_shell_problem(refinement, N, radial_partitioning=layers)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ctx, part = ev.ctx, ev.part
    assert part.oriented and len(part.external_faces) == (problem.neighbors == -1).sum()
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    rng = np.random.default_rng(N)
    u = u0 + 1e-3 * rng.uniform(-1, 1, u0.shape)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    H, dH = _gauge_fields(N, x, J, u0)
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    ref = orc.dg_rhs(1, N, u, J, np.concatenate([stat, H, dH], axis=1), part.local_neighbors,
                     gauge_params=orc.GAUGE_GIVEN, ext_u=ext,
                     nbr_dir=part.local_neighbor_direction,
                     face_perm=part.local_face_permutation)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    # ignoring the orientations gives a different answer: they matter
    wrong = orc.dg_rhs(1, N, u, J, np.concatenate([stat, H, dH], axis=1), part.local_neighbors,
                       gauge_params=orc.GAUGE_GIVEN, ext_u=ext)
    assert _relerr(wrong, ref, GH_BLOCKS) > 1e-6
    ctx.close()


def test_scalar_wave_on_shell_matches_oracle():
    """ScalarWave on the same six-wedge shell (outflow on both boundaries)."""
    N = 6
    sh = domain.SphericalShell(1.0, 3.0, (1, 0), N, radial_distribution="Linear")
    x, J, nbr = sh.coords(), sh.inverse_jacobian(), sh.neighbors()
    nd, perm = sh.neighbor_orientations()
    rng = np.random.default_rng(3)
    u = analytic.plane_wave(x, 0.2) + 0.05 * rng.uniform(-1, 1, (sh.n_elements, 5, N ** 3))
    stat = rng.uniform(0, 1, (sh.n_elements, 1, N ** 3))
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, sh.n_elements)
    ctx.set_geometry(J, x, nbr)
    ctx.set_neighbor_orientations(nd, perm)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(0, N, u, J, stat, nbr, nbr_dir=nd, face_perm=perm)
    assert _relerr(got, ref, SW_BLOCKS) < TOL
    dt = 1e-3
    ctx.set_stepper(lib.STEPPER_RK3_HESTHAVEN, 3, 0.0, dt)
    ctx.take_steps(2)
    ev = orc.Evolution(lambda v, t: orc.dg_rhs(0, N, v, J, stat, nbr, nbr_dir=nd,
                                               face_perm=perm), u, 0.0, dt, "RK3")
    ev.step()
    ev.step()
    assert _relerr(ctx.get_state(), ev.u, SW_BLOCKS) < TOL
    ctx.close()


def test_kerr_schild_yaml_configuration():
    """tests/InputFiles/GeneralizedHarmonic/KerrSchild.yaml: Sphere r in
    [1.9, 2.3], InitialRefinement 0, InitialGridPoints 5, Logarithmic, equiangular,
    DirichletAnalytic on both boundaries, AnalyticChristoffel gauge,
    AdamsBashforth order 4, step 2e-4, exponential filter (Alpha 36, HalfPower
    64, :127-132).  GPU evolution vs the oracle over the self-start and 4 steps;
    the exact static solution is kept to truncation level."""
    N, dt = 5, 2e-4
    problem = evolution.gh_kerr_schild_shell_problem((0, 0), N)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 4, dt)
    ctx, part = ev.ctx, ev.part
    assert part.n_local == 6
    ctx.set_exponential_filter(True, 36.0, 64)
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    H, dH = _gauge_fields(N, x, J, u0)
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    sf = np.concatenate([stat, H, dH], axis=1)
    F = orc.exponential_filter_matrix(N, 36.0, 64)

    def rhs(v, t):
        return orc.dg_rhs(1, N, v, J, sf, part.local_neighbors, gauge_params=orc.GAUGE_GIVEN,
                          ext_u=ext, nbr_dir=part.local_neighbor_direction,
                          face_perm=part.local_face_permutation)
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 4, 0.0, dt)
    ev.take_steps(4)
    oev = orc.Evolution(rhs, u0, 0.0, dt, "AB4", post_update=lambda v: orc.apply_filter(N, v, F))
    for _ in range(4):
        oev.step()
    got = ctx.get_state()
    assert ctx.rhs_evaluations == oev.rhs_evals
    assert _relerr(got, oev.u, GH_BLOCKS) < TOL
    # Error(SpacetimeMetric, Pi, Phi) as observed by KerrSchild.yaml:153-165: at this
    # resolution (5 points across a 90 degree wedge) it is the top Legendre mode that
    # the filter removes from the analytic data, not a drift of the evolution
    top_mode = np.max(np.abs(orc.apply_filter(N, u0, F) - u0))
    assert 1e-3 < top_mode < 0.2
    assert np.max(np.abs(got - u0)) < 1.5 * top_mode
    # constraint norms (ObserveNorms of the constraint tags) agree with the oracle's
    cg = ctx.gh_constraint_norms()
    co = orc.gh_constraint_norms(N, got, J, H)
    np.testing.assert_allclose(cg, co, rtol=1e-9, atol=1e-14)
    ctx.close()


@pytest.mark.parametrize("N,order", [(10, 3), (12, 3)])
def test_kerr_schild_shell_evolution_north_star_points(N, order):
    """BASELINE configs[2]/[3] at their own numbers of grid points (P = 9 / P = 11):
    Kerr-Schild shell, DirichletAnalytic on both spheres, AnalyticChristoffel gauge
    (the gauge-fields kernel), AB3 with the exponential filter of KerrSchild.yaml
    :127-132; self-start + 5 steps against the oracle, evolved tensors and the
    three constraint norms."""
    dt = 1e-4
    problem = evolution.gh_kerr_schild_shell_problem((0, 0), N)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, order, dt)
    ctx, part = ev.ctx, ev.part
    ctx.set_exponential_filter(True, 36.0, 64)
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    H, dH = _gauge_fields(N, x, J, u0)
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    sf = np.concatenate([stat, H, dH], axis=1)
    F = orc.exponential_filter_matrix(N, 36.0, 64)

    def rhs(v, t):
        return orc.dg_rhs(1, N, v, J, sf, part.local_neighbors, gauge_params=orc.GAUGE_GIVEN,
                          ext_u=ext, nbr_dir=part.local_neighbor_direction,
                          face_perm=part.local_face_permutation)
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, order, 0.0, dt)
    ev.take_steps(5)
    oev = orc.Evolution(rhs, u0, 0.0, dt, f"AB{order}",
                        post_update=lambda v: orc.apply_filter(N, v, F))
    for _ in range(5):
        oev.step()
    got = ctx.get_state()
    assert ctx.rhs_evaluations == oev.rhs_evals
    assert _relerr(got, oev.u, GH_BLOCKS) < TOL
    # the exact static solution is kept up to the top Legendre mode the filter removes
    top_mode = np.max(np.abs(orc.apply_filter(N, u0, F) - u0))
    assert np.max(np.abs(got - u0)) < max(1.5 * top_mode, 1e-9)
    cg = ctx.gh_constraint_norms()
    co = orc.gh_constraint_norms(N, got, J, H)
    np.testing.assert_allclose(cg, co, rtol=1e-8, atol=1e-13)
    ctx.close()


def test_demand_outgoing_char_speeds_on_excision_boundary():
    """DemandOutgoingCharSpeeds on the excision sphere (inside the horizon all
    characteristic fields leave the domain): no correction on those faces (the
    oracle sees neighbour -1 there), the check passes for r = 1.9 and reports
    the violation for an 'excision' sphere outside the horizon."""
    N = 5
    problem = evolution.gh_kerr_schild_shell_problem((0, 0), N,
                                                     inner_boundary="DemandOutgoingCharSpeeds")
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ctx, part = ev.ctx, ev.part
    assert len(part.external_faces) == 6 and (part.local_neighbors == -1).sum() == 6
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    u = u0 + 1e-3 * np.random.default_rng(8).uniform(-1, 1, u0.shape)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    H, dH = _gauge_fields(N, x, J, u0)
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    ref = orc.dg_rhs(1, N, u, J, np.concatenate([stat, H, dH], axis=1), part.local_neighbors,
                     gauge_params=orc.GAUGE_GIVEN, ext_u=ext,
                     nbr_dir=part.local_neighbor_direction,
                     face_perm=part.local_face_permutation)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    ctx.check_outgoing_char_speeds()      # all four speeds >= 0 on r = 1.9 < 2M
    ev.take_steps(2)
    ctx.check_outgoing_char_speeds()
    ctx.close()
    # outside the horizon lambda_- = -beta.n - alpha < 0 with respect to the outward
    # (towards the hole) normal: the reference ERRORs, the library reports it
    bad = evolution.gh_kerr_schild_shell_problem((0, 0), N, inner_radius=2.6, outer_radius=3.0,
                                                 inner_boundary="DemandOutgoingCharSpeeds")
    ev = evolution.Evolution(bad, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ev.ctx.compute_time_derivative(0.0)
    with pytest.raises(lib.DgrhsError, match="DemandOutgoingCharSpeeds boundary condition "
                                             "violated"):
        ev.ctx.check_outgoing_char_speeds()
    # number of violating face points and the worst speed agree with the oracle
    ids = ev.part.global_ids
    n_bad, worst = orc.demand_outgoing_char_speeds(N, bad.u0(ids, 0.0),
                                                   bad.inverse_jacobian(ids),
                                                   bad.static(ids)[:, 1], ev.part.local_neighbors)
    assert n_bad == 6 * N * N
    n = ctypes.c_longlong(0)
    mn = ctypes.c_double(0.0)
    rc = lib.load().dgrhs_check_outgoing_char_speeds(ev.ctx._h, ctypes.byref(n), ctypes.byref(mn))
    assert rc != 0 and n.value == n_bad
    assert mn.value == pytest.approx(worst, rel=1e-12)
    ev.ctx.close()


@pytest.mark.parametrize("outer", ["DirichletAnalytic", "ConstraintPreservingPhysical"])
def test_partitioned_shell_matches_single_context(outer):
    """The multi-GPU path on the multi-block shell (ghost faces with non-aligned
    orientation + DirichletAnalytic slots or Bjorhus faces, which also sit on
    "interior" elements of a rank), run as 3 contexts on one GPU with the halo
    moved by device copies: bit-identical to the single-context evolution."""
    import torch
    N, dt, world = 4, 1e-4, 3
    problem = evolution.gh_kerr_schild_shell_problem((1, 0), N, outer_boundary=outer)
    single = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, dt)
    single.take_steps(2)
    ref = single.gather_state(problem.brick.n_elements)
    single.ctx.close()

    class _FakeDist:  # Evolution only stores it for world > 1
        pass
    evs = []
    for r in range(world):
        ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, dt, device=0,
                                 world=world, rank=r)
        evs.append(ev)
    per_face = evs[0].ctx.halo_comps * N * N

    def offsets(counts):
        out, o = [], 0
        for cnt in counts:
            out.append(o)
            o += cnt
        return out
    done = 0
    while done < 2:
        times = [ev.ctx.begin_substep() for ev in evs]
        assert len(set(times)) == 1
        for ev in evs:
            ev.ctx.pack_halo()
            ev.ctx.synchronize()
        for r in range(world):
            ro = offsets(evs[r].part.recv_counts)
            for p in range(world):
                cnt = evs[r].part.recv_counts[p]
                if cnt == 0:
                    continue
                so = offsets(evs[p].part.send_counts)[r]
                evs[r]._recv[ro[p] * per_face:(ro[p] + cnt) * per_face].copy_(
                    evs[p]._send[so * per_face:(so + cnt) * per_face])
        torch.cuda.synchronize()
        fin = []
        for ev in evs:
            ev.ctx.compute_time_derivative_range(times[0], 0, ev.part.n_interior)
            ev.ctx.compute_time_derivative_range(times[0], ev.part.n_interior, ev.part.n_local)
            fin.append(ev.ctx.end_substep())
        assert len(set(fin)) == 1
        done += int(fin[0])
    out = np.empty_like(ref)
    for ev in evs:
        out[ev.part.global_ids] = ev.ctx.get_state()
        ev.ctx.close()
    np.testing.assert_array_equal(out, ref)


def test_shell_with_layers_of_different_refinement():
    """Sphere with per-layer InitialRefinement (inner layer one angular level
    finer): the spherical interface between the layers consists of non-conforming
    2:1 mortars, next to non-aligned conforming neighbours between the wedges and
    DirichletAnalytic ghosts on both boundaries -- all face kinds in one RHS."""
    N = 5
    problem = evolution.gh_kerr_schild_shell_problem([(2, 0), (1, 0)], N,
                                                     radial_partitioning=(2.1,))
    assert len(problem.mortars) == 6 * 16
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ctx, part = ev.ctx, ev.part
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    u = u0 + 1e-3 * np.random.default_rng(4).uniform(-1, 1, u0.shape)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    H, dH = _gauge_fields(N, x, J, u0)
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    sf = np.concatenate([stat, H, dH], axis=1)
    kw = dict(gauge_params=orc.GAUGE_GIVEN, ext_u=ext, nbr_dir=part.local_neighbor_direction,
              face_perm=part.local_face_permutation)
    ref = orc.dg_rhs(1, N, u, J, sf, part.local_neighbors, mortars=ev.local_mortars, **kw)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    without = orc.dg_rhs(1, N, u, J, sf,
                         np.where(part.local_neighbors == domain.HANGING, -1,
                                  part.local_neighbors), **kw)
    assert _relerr(without, ref, GH_BLOCKS) > 1e-6
    # the static solution stays put on the non-conforming shell
    ctx.set_state(u0)
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, 1e-4)
    ev.take_steps(3)
    assert np.max(np.abs(ctx.get_state() - u0)) < 1e-4
    ctx.close()


@pytest.mark.parametrize("fine", [{0: (1, 1)}, {5: (1, 1), 4: (1, 0)}, {2: (1, 0), 3: (1, 0)}])
def test_shell_with_wedges_of_different_refinement(fine):
    """Sphere with per-block InitialRefinement: refined wedges next to coarse ones, so
    the wedge-to-wedge interfaces are 2:1 mortars (four quarters, or two halves when
    only the angular level differs) between blocks that are NOT aligned -- mortar rows
    with face permutations, next to oriented conforming faces and ghost boundaries."""
    N = 5
    ref = [fine.get(w, (0, 0)) for w in range(6)]
    problem = evolution.gh_kerr_schild_shell_problem([ref], N)
    mt = np.asarray(problem.mortars)
    assert len(mt) > 0 and ((mt[:, 3] >> 3) != 0).any()
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ctx, part = ev.ctx, ev.part
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    u = u0 + 1e-3 * np.random.default_rng(4).uniform(-1, 1, u0.shape)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    H, dH = _gauge_fields(N, x, J, u0)
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    sf = np.concatenate([stat, H, dH], axis=1)
    kw = dict(gauge_params=orc.GAUGE_GIVEN, ext_u=ext, nbr_dir=part.local_neighbor_direction,
              face_perm=part.local_face_permutation)
    ref_rhs = orc.dg_rhs(1, N, u, J, sf, part.local_neighbors, mortars=ev.local_mortars, **kw)
    assert _relerr(got, ref_rhs, GH_BLOCKS) < TOL
    stripped = np.array(ev.local_mortars).copy()
    stripped[:, 3] &= 7
    wrong = orc.dg_rhs(1, N, u, J, sf, part.local_neighbors, mortars=stripped, **kw)
    assert _relerr(wrong, ref_rhs, GH_BLOCKS) > 1e-3
    # the static solution stays put
    ctx.set_state(u0)
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, 1e-4)
    ev.take_steps(3)
    assert np.max(np.abs(ctx.get_state() - u0)) < 1e-3
    ctx.close()
