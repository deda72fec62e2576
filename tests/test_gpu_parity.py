"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU
oracle on the same seeded inputs.  Tolerance: 1e-12 relative in the max norm
over each evolved tensor (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, lib
from tests.test_oracle_pins import _random_physical_gh_state

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _relerr(got, ref, blocks):
    return max(np.max(np.abs(got[:, b] - ref[:, b])) / np.max(np.abs(ref[:, b])) for b in blocks)


SW_BLOCKS = [slice(0, 1), slice(1, 2), slice(2, 5)]
GH_BLOCKS = [slice(0, 10), slice(10, 20), slice(20, 50)]


def _curved_jacobian(rng, brick):
    """Full 3x3 inverse Jacobian with off-diagonal terms (a smooth, per-point
    perturbation of the affine one) so that every J entry is exercised."""
    J = brick.inverse_jacobian()
    J = J + 0.1 * rng.uniform(-1, 1, J.shape)
    return J


@pytest.mark.parametrize("N", [2, 3, 4, 5, 6, 7, 8])
def test_partial_derivatives_operator(N):
    rng = np.random.default_rng(N)
    u = rng.uniform(-1, 1, (7, N ** 3))
    J = rng.uniform(-1, 1, (9, N ** 3))
    got = lib.partial_derivatives(N, u, J)
    ref = orc.partial_derivatives(N, u, J)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < TOL


@pytest.mark.parametrize("N,refine", [(2, 1), (3, 1), (5, 1), (6, 2), (7, 1), (8, 1), (9, 1),
                                      (10, 1), (11, 0), (12, 1)])
def test_scalar_wave_rhs(N, refine):
    rng = np.random.default_rng(100 + N)
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [refine] * 3, N)
    x = brick.coords()
    u = analytic.plane_wave(x, 0.3) + 0.1 * rng.uniform(-1, 1, (brick.n_elements, 5, brick.n))
    J = _curved_jacobian(rng, brick)
    nb = brick.neighbors()
    stat = rng.uniform(0, 1, (brick.n_elements, 1, brick.n))
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    for volume_only in (True, False):
        ctx.compute_time_derivative(0.0, volume_only=volume_only)
        got = ctx.get_time_derivative()
        ref = orc.dg_rhs(0, N, u, J, stat, nb, volume_only=volume_only)
        assert _relerr(got, ref, SW_BLOCKS) < TOL
    np.testing.assert_array_equal(ctx.get_state(), u)  # round trip, bit exact
    ctx.close()


def _gh_problem(rng, N, refine, noise=1e-2):
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [refine] * 3, N)
    x = brick.coords()
    u = analytic.gauge_wave(x, 0.1)
    u = u + noise * rng.uniform(-1, 1, u.shape)
    J = _curved_jacobian(rng, brick)
    stat = rng.uniform(-1, 1, (brick.n_elements, 3, brick.n))
    return brick, x, u, J, stat


@pytest.mark.parametrize("N,refine", [(2, 1), (3, 1), (4, 1), (5, 1), (6, 1), (7, 1), (8, 1),
                                      (9, 0), (10, 1), (11, 0), (12, 1)])
def test_gh_rhs_harmonic(N, refine):
    rng = np.random.default_rng(200 + N)
    brick, x, u, J, stat = _gh_problem(rng, N, refine)
    nb = brick.neighbors()
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    for volume_only in (True, False):
        ctx.compute_time_derivative(0.0, volume_only=volume_only)
        got = ctx.get_time_derivative()
        ref = orc.dg_rhs(1, N, u, J, stat, nb, volume_only=volume_only)
        assert _relerr(got, ref, GH_BLOCKS) < TOL
    ctx.close()


def test_gh_rhs_random_physical_state():
    """Far-from-flat random physical metrics (like Test_DuDt.cpp:489-493), shift
    large enough that characteristic speeds change sign on the faces."""
    N = 5
    rng = np.random.default_rng(77)
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    u = np.stack([_random_physical_gh_state(rng, brick.n) for _ in range(brick.n_elements)])
    J = _curved_jacobian(rng, brick)
    stat = rng.uniform(-1.5, 1, (brick.n_elements, 3, brick.n))
    nb = brick.neighbors()
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, None, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(1, N, u, J, stat, nb)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    ctx.close()


@pytest.mark.parametrize("N,refine", [(6, 1), (8, 1), (9, 0), (10, 1), (11, 0), (12, 1)])
def test_gh_rhs_gauge_fields(N, refine):
    """The gauge-fields instantiation of the volume kernel (H_a, d_a H_b from memory:
    the AnalyticChristoffel gauge of the Kerr-Schild configs) at every N the
    Kerr-Schild configs use; N >= 10 runs 256-point chunks, one CTA per SM."""
    rng = np.random.default_rng(9 + N)
    brick, x, u, J, stat = _gh_problem(rng, N, refine)
    nb = brick.neighbors()
    H = rng.uniform(-1, 1, (brick.n_elements, 4, brick.n))
    dH = rng.uniform(-1, 1, (brick.n_elements, 16, brick.n))
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_gauge(lib.GAUGE_FIELDS)
    ctx.set_gauge_fields(H, dH)
    ctx.set_state(u)
    for volume_only in (True, False):
        ctx.compute_time_derivative(0.0, volume_only=volume_only)
        got = ctx.get_time_derivative()
        ref = orc.dg_rhs(1, N, u, J, np.concatenate([stat, H, dH], axis=1), nb,
                         gauge_params=orc.GAUGE_GIVEN, volume_only=volume_only)
        assert _relerr(got, ref, GH_BLOCKS) < TOL
    ctx.close()


def test_external_boundaries_get_no_correction():
    N = 4
    rng = np.random.default_rng(4)
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N, periodic=(True, False, True))
    x = brick.coords()
    u = analytic.plane_wave(x, 0.0)
    J = brick.inverse_jacobian()
    stat = rng.uniform(0, 1, (brick.n_elements, 1, brick.n))
    nb = brick.neighbors()
    assert (nb == -1).any()
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    ref = orc.dg_rhs(0, N, u, J, stat, nb)
    assert _relerr(ctx.get_time_derivative(), ref, SW_BLOCKS) < TOL
    ctx.close()


@pytest.mark.parametrize("stepper", ["AB1", "AB2", "AB3", "AB4", "RK3"])
def test_scalar_wave_evolution(stepper):
    """Config 1 at reduced size (2^3 elements, N = 6): self-start + 4 steps."""
    N, dt = 6, 1e-3
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    u0 = analytic.plane_wave(x, 0.0)
    stat = np.zeros((brick.n_elements, 1, brick.n))
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u0)
    if stepper == "RK3":
        ctx.set_stepper(lib.STEPPER_RK3_HESTHAVEN, 3, 0.0, dt)
    else:
        ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, int(stepper[2:]), 0.0, dt)
    ctx.take_steps(4)
    ev = orc.Evolution(lambda u, t: orc.dg_rhs(0, N, u, J, stat, nb), u0, 0.0, dt, stepper)
    for _ in range(4):
        ev.step()
    got = ctx.get_state()
    assert ctx.rhs_evaluations == ev.rhs_evals
    assert abs(ctx.time - ev.time) < 1e-15
    assert _relerr(got, ev.u, SW_BLOCKS) < TOL
    # exact-solution error norm (PlaneWave3D.yaml observes Error(...) norms)
    exact = analytic.plane_wave(x, ctx.time)
    e_gpu, e_cpu = orc.l2_norm(got - exact), orc.l2_norm(ev.u - exact)
    assert e_gpu < 1e-3 and abs(e_gpu - e_cpu) < 1e-9 * e_cpu + 1e-13
    ctx.close()


def test_gh_gauge_wave_evolution_ab3():
    """GaugeWave3D.yaml at its CI size (2^3 elements, N = 5, AB3, dt = 2e-4,
    gamma0 = 1, gamma1 = -1, gamma2 = 1): self-start + 5 steps; evolved tensors
    and the error norms against the exact solution."""
    N, dt = 5, 2e-4
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    u0 = analytic.gauge_wave(x, 0.0)
    stat = np.zeros((brick.n_elements, 3, brick.n))
    stat[:, 0], stat[:, 1], stat[:, 2] = 1.0, -1.0, 1.0
    for gauge in (lib.GAUGE_HARMONIC, lib.GAUGE_ANALYTIC_GAUGE_WAVE):
        ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
        ctx.set_geometry(J, x, nb)
        ctx.set_static_fields(stat)
        if gauge == lib.GAUGE_ANALYTIC_GAUGE_WAVE:
            ctx.set_gauge(gauge, [0.1, 1.0])
        ctx.set_state(u0)
        ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
        ctx.take_steps(5)

        def rhs(u, t):
            if gauge == lib.GAUGE_HARMONIC:
                return orc.dg_rhs(1, N, u, J, stat, nb)
            H = np.zeros((brick.n_elements, 4, brick.n))
            dH = np.zeros((brick.n_elements, 16, brick.n))
            for e in range(brick.n_elements):
                ua = orc.gh_vars_from_metric(*orc.gauge_wave_metric(x[e], t))
                H[e], dH[e] = orc.analytic_christoffel_gauge(N, ua, J[e])
            return orc.dg_rhs(1, N, u, J, np.concatenate([stat, H, dH], axis=1), nb,
                              gauge_params=orc.GAUGE_GIVEN)

        ev = orc.Evolution(rhs, u0, 0.0, dt, "AB3")
        for _ in range(5):
            ev.step()
        got = ctx.get_state()
        assert _relerr(got, ev.u, GH_BLOCKS) < TOL
        exact = analytic.gauge_wave(x, ctx.time)
        for b in GH_BLOCKS:
            e_gpu = orc.l2_norm(got[:, b] - exact[:, b])
            e_cpu = orc.l2_norm(ev.u[:, b] - exact[:, b])
            assert abs(e_gpu - e_cpu) <= 1e-12 * max(1.0, e_cpu) + 1e-9 * e_cpu
        ctx.close()


def _wrap(ptr, count):
    import torch
    from spectre_b200.evolution import _CudaArray
    return torch.as_tensor(_CudaArray(ptr, count), device="cuda:0")


@pytest.mark.parametrize("system,world", [("gh", 2), ("sw", 2), ("gh", 4)])
def test_partitioned_evolution_matches_single_context(system, world):
    """The multi-GPU path (Partition, pack_halo, ghost faces in the face kernel,
    interior/boundary ranges, substep API) run as `world` contexts on ONE GPU
    with the halo moved by device copies instead of NCCL: the gathered state
    must be bit-identical to the single-context evolution."""
    from spectre_b200 import evolution
    N = 4
    if system == "gh":
        problem = evolution.gh_gauge_wave_problem([2, 1, 1], N)
        dt = 2e-4
    else:
        problem = evolution.scalar_wave_problem([2, 1, 1], N)
        dt = 1e-3
    single = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, dt)
    single.take_steps(3)
    ref = single.gather_state(problem.brick.n_elements)

    evs = [evolution.Evolution.__new__(evolution.Evolution) for _ in range(world)]
    parts = [domain.Partition(problem.neighbors, world, r) for r in range(world)]
    ctxs = []
    f = N * N
    for r, part in enumerate(parts):
        ids = part.global_ids
        ctx = lib.Context(problem.system, N, part.n_local, part.n_ghost, 0)
        ctx.set_geometry(problem.inverse_jacobian(ids), problem.coords(ids),
                         part.local_neighbors)
        ctx.set_static_fields(problem.static(ids))
        ctx.set_state(problem.u0(ids, 0.0))
        ctx.set_halo_map(part.send_map)
        ctx.set_interior_count(part.n_interior)
        ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
        ctxs.append(ctx)
    per_face = ctxs[0].halo_comps * f
    send = [_wrap(c.halo_send_ptr(), p.n_ghost * per_face) for c, p in zip(ctxs, parts)]
    recv = [_wrap(c.halo_recv_ptr(), p.n_ghost * per_face) for c, p in zip(ctxs, parts)]

    def offsets(counts):
        out, o = [], 0
        for cnt in counts:
            out.append(o)
            o += cnt
        return out

    steps_done = 0
    while steps_done < 3:
        times = [c.begin_substep() for c in ctxs]
        assert len(set(times)) == 1
        for c in ctxs:
            c.pack_halo()
            c.synchronize()
        for r in range(world):          # receiver
            ro = offsets(parts[r].recv_counts)
            for p in range(world):      # sender
                cnt = parts[r].recv_counts[p]
                if cnt == 0:
                    continue
                assert parts[p].send_counts[r] == cnt
                so = offsets(parts[p].send_counts)[r]
                recv[r][ro[p] * per_face:(ro[p] + cnt) * per_face].copy_(
                    send[p][so * per_face:(so + cnt) * per_face])
        import torch
        torch.cuda.synchronize()
        done = []
        for c, part in zip(ctxs, parts):
            if part.n_interior > 0:
                c.compute_time_derivative_range(times[0], 0, part.n_interior)
            c.compute_time_derivative_range(times[0], part.n_interior, part.n_local)
            done.append(c.end_substep())
        assert len(set(done)) == 1
        steps_done += int(done[0])
    got = np.empty_like(ref)
    for c, part in zip(ctxs, parts):
        got[part.global_ids] = c.get_state()
        assert c.rhs_evaluations == single.ctx.rhs_evaluations
    np.testing.assert_array_equal(got, ref)


def test_gh_rhs_damped_harmonic():
    """DampedHarmonic gauge (DampedHarmonic.cpp:70-439) inside the fused volume
    kernel vs the oracle (pinned by the reference's DampedHarmonic.py); the
    parameters are those of Test_DuDt.cpp's DampedHarmonic{100, {1.2, 1.5, 1.7},
    {2, 4, 6}} with a smaller width so that the spatial weight varies."""
    N = 6
    rng = np.random.default_rng(31)
    brick, x, u, J, stat = _gh_problem(rng, N, 1, noise=5e-2)
    nb = brick.neighbors()
    params = [3.0, 1.2, 1.5, 1.7, 2, 4, 6]
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_gauge(lib.GAUGE_DAMPED_HARMONIC, params)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(1, N, u, J, stat, nb, gauge_params=np.array([2.0] + params), coords=x)
    ref_h = orc.dg_rhs(1, N, u, J, stat, nb)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    assert _relerr(ref_h, ref, GH_BLOCKS) > 1e-6  # the gauge terms do matter here
    ctx.close()


def test_gh_constraint_norms():
    """GH constraint norms (north_star parity metric): device diagnostics vs the
    oracle on a perturbed gauge wave, and after a short evolution the norms of
    the GPU state equal those of the oracle state to 1e-12 relative."""
    N, dt = 5, 2e-4
    rng = np.random.default_rng(12)
    brick, x, u, J, stat = _gh_problem(rng, N, 1, noise=1e-3)
    J = brick.inverse_jacobian()
    nb = brick.neighbors()
    stat[:, 0], stat[:, 1], stat[:, 2] = 1.0, -1.0, 1.0
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    got = ctx.gh_constraint_norms()
    ref = orc.gh_constraint_norms(N, u, J)
    np.testing.assert_allclose(got, ref, rtol=1e-11)
    assert (ref > 1e-6).all()
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
    ctx.take_steps(3)
    ev = orc.Evolution(lambda v, t: orc.dg_rhs(1, N, v, J, stat, nb), u, 0.0, dt, "AB3")
    for _ in range(3):
        ev.step()
    np.testing.assert_allclose(ctx.gh_constraint_norms(), orc.gh_constraint_norms(N, ev.u, J),
                               rtol=1e-11)
    np.testing.assert_allclose(orc.gh_constraint_norms(N, ctx.get_state(), J),
                               orc.gh_constraint_norms(N, ev.u, J), rtol=1e-12)
    ctx.close()


@pytest.mark.parametrize("system,stepper", [("gh", "AB3"), ("gh", "RK3"), ("sw", "AB4"),
                                            ("sw", "RK3"), ("gh", "AB1")])
def test_fused_update_is_bit_identical_to_separate_update(system, stepper):
    """UpdateU fused into the volume kernel (default) vs the separate
    lincomb_kernel: same coefficients, same term order -> identical bits."""
    from spectre_b200 import evolution
    N = 5
    problem = (evolution.gh_gauge_wave_problem([1, 1, 1], N) if system == "gh"
               else evolution.scalar_wave_problem([1, 1, 1], N))
    dt = 2e-4 if system == "gh" else 1e-3
    states = []
    for fused in (True, False):
        if stepper == "RK3":
            ev = evolution.Evolution(problem, lib.STEPPER_RK3_HESTHAVEN, 3, dt)
        else:
            ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, int(stepper[2:]), dt)
        ev.ctx.set_fused_update(fused)
        ev.ctx.take_steps(4)
        states.append(ev.ctx.get_state())
        ev.ctx.close()
    np.testing.assert_array_equal(states[0], states[1])


def test_gh_kerr_schild_dirichlet_analytic():
    """Kerr-Schild on a non-periodic Brick with DirichletAnalytic on every
    external face (ghost boundary condition, BoundaryConditionsImpl.hpp:427-560
    + DirichletAnalytic.cpp:58-117), AnalyticChristoffel gauge of the static
    solution and the GaussianPlusConstant damping functions of KerrSchild.yaml."""
    from spectre_b200 import evolution
    N, dt = 6, 1e-3
    problem = evolution.gh_kerr_schild_problem([1, 1, 1], N)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, dt)
    ctx, part = ev.ctx, ev.part
    assert len(part.external_faces) == 24 and part.n_recv == 0
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    # perturb the state so that the boundary correction is not trivially small
    rng = np.random.default_rng(2)
    u = u0 + 1e-3 * rng.uniform(-1, 1, u0.shape)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    H = np.zeros((len(ids), 4, N ** 3)); dH = np.zeros((len(ids), 16, N ** 3))
    for e in range(len(ids)):
        ua = orc.gh_vars_from_metric(*orc.kerr_schild_metric(x[e]))
        np.testing.assert_allclose(ua, u0[e], rtol=1e-13, atol=1e-14)
        H[e], dH[e] = orc.analytic_christoffel_gauge(N, ua, J[e])
    ext = ev.boundary_ghost_data(problem, 0.0)[:, :50]
    ref = orc.dg_rhs(1, N, u, J, np.concatenate([stat, H, dH], axis=1), part.local_neighbors,
                     gauge_params=orc.GAUGE_GIVEN, ext_u=ext)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    # the same without the boundary condition differs: the BC matters
    ref_nobc = orc.dg_rhs(1, N, u, J, np.concatenate([stat, H, dH], axis=1),
                          np.where(part.local_neighbors < -1, -1, part.local_neighbors),
                          gauge_params=orc.GAUGE_GIVEN)
    assert _relerr(ref_nobc, ref, GH_BLOCKS) > 1e-4
    # the exact (static) solution is kept: |dt u| is at truncation level and the
    # evolved state stays at the analytic one
    ctx.set_state(u0)
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
    ctx.take_steps(5)
    drift = np.max(np.abs(ctx.get_state() - u0))
    assert drift < 1e-6, drift
    ctx.close()


def test_exponential_filter_evolution():
    """dg::Actions::Filter<Exponential<0>> after every substep update: GPU vs
    oracle over a self-started AB3 evolution of a perturbed gauge wave with a
    filter that visibly damps (alpha = 4, half power = 2)."""
    N, dt = 5, 2e-4
    rng = np.random.default_rng(17)
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    u0 = analytic.gauge_wave(x, 0.0) + 1e-3 * rng.uniform(-1, 1, (brick.n_elements, 50, N ** 3))
    stat = np.zeros((brick.n_elements, 3, brick.n))
    stat[:, 0], stat[:, 1], stat[:, 2] = 1.0, -1.0, 1.0
    F = orc.exponential_filter_matrix(N, 4.0, 2)
    results = []
    for stepper in ("AB3", "RK3"):
        ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
        ctx.set_geometry(J, x, nb)
        ctx.set_static_fields(stat)
        ctx.set_exponential_filter(True, 4.0, 2)
        ctx.set_state(u0)
        if stepper == "RK3":
            ctx.set_stepper(lib.STEPPER_RK3_HESTHAVEN, 3, 0.0, dt)
        else:
            ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
        ctx.take_steps(3)
        ev = orc.Evolution(lambda v, t: orc.dg_rhs(1, N, v, J, stat, nb), u0, 0.0, dt, stepper,
                           post_update=lambda v: orc.apply_filter(N, v, F))
        for _ in range(3):
            ev.step()
        got = ctx.get_state()
        assert _relerr(got, ev.u, GH_BLOCKS) < TOL
        results.append(got)
        ctx.close()
    # the filter did something: unfiltered evolution differs
    ev = orc.Evolution(lambda v, t: orc.dg_rhs(1, N, v, J, stat, nb), u0, 0.0, dt, "AB3")
    for _ in range(3):
        ev.step()
    assert np.max(np.abs(ev.u - results[0])) > 1e-6


def test_config1_full_size_scalar_wave_rk3():
    """BASELINE.json configs[0] at full size: ScalarWave plane wave, periodic
    Brick [0,2pi]^3, 8^3 elements, N = 6, Rk3HesthavenSsp, dt = 1e-3."""
    from spectre_b200 import evolution
    N, dt = 6, 1e-3
    problem = evolution.scalar_wave_problem([3, 3, 3], N)
    ev = evolution.Evolution(problem, lib.STEPPER_RK3_HESTHAVEN, 3, dt)
    ev.take_steps(3)
    got = ev.ctx.get_state()
    x, J = problem.coords(), problem.inverse_jacobian()
    stat, nb = problem.static(), problem.neighbors
    o = orc.Evolution(lambda u, t: orc.dg_rhs(0, N, u, J, stat, nb), problem.u0(), 0.0, dt, "RK3")
    for _ in range(3):
        o.step()
    assert _relerr(got, o.u, SW_BLOCKS) < TOL
    exact = analytic.plane_wave(x, ev.ctx.time)
    assert orc.l2_norm(got - exact) < 1e-6
    ev.ctx.close()


def test_config2_full_size_gauge_wave():
    """BASELINE.json configs[1] at full size (16^3 elements, N = 8): one RHS
    against the oracle, then size-independent properties of a 5-step AB3
    evolution: the error against the exact gauge wave stays at truncation level,
    the solution is invariant under translation by one element in y and z
    (the data depend on x only), and the constraint norms stay small."""
    from spectre_b200 import evolution
    N, dt = 8, 2e-4
    problem = evolution.gh_gauge_wave_problem([4, 4, 4], N)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, dt)
    ctx = ev.ctx
    u0 = problem.u0()
    J, stat, nb = problem.inverse_jacobian(), problem.static(), problem.neighbors
    ctx.compute_time_derivative(0.0)
    ref = orc.dg_rhs(1, N, u0, J, stat, nb)
    assert _relerr(ctx.get_time_derivative(), ref, GH_BLOCKS) < TOL
    ev.take_steps(5)
    got = ctx.get_state()
    exact = analytic.gauge_wave(problem.coords(), ctx.time)
    assert np.max(np.abs(got - exact)) < 1e-9
    cells = problem.brick.cells
    index_of = problem.brick.index_of
    shifted = np.array([index_of[(c[0], (c[1] + 1) % 16, (c[2] + 3) % 16)] for c in cells])
    assert np.max(np.abs(got - got[shifted])) < 1e-13
    assert (ctx.gh_constraint_norms() < 1e-8).all()
    ctx.close()


def test_time_dependent_dirichlet_analytic_gauge_wave():
    """DirichletAnalytic with a time-dependent solution: the exterior state is
    re-evaluated at the time of every RHS (incl. the self-start substeps)."""
    from spectre_b200 import evolution
    N, dt = 5, 2e-4
    problem = evolution.gh_gauge_wave_dirichlet_problem([1, 1, 1], N)
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, dt)
    part = ev.part
    assert len(part.external_faces) == 8
    ev.take_steps(4)
    got = ev.ctx.get_state()
    ids = part.global_ids
    J, stat = problem.inverse_jacobian(ids), problem.static(ids)

    def rhs(u, t):
        ext = ev.boundary_ghost_data(problem, t)[:, :50]
        return orc.dg_rhs(1, N, u, J, stat, part.local_neighbors, ext_u=ext)

    o = orc.Evolution(rhs, problem.u0(ids, 0.0), 0.0, dt, "AB3")
    for _ in range(4):
        o.step()
    assert _relerr(got, o.u, GH_BLOCKS) < TOL
    exact = problem.u0(ids, ev.ctx.time)
    e_gpu, e_cpu = np.max(np.abs(got - exact)), np.max(np.abs(o.u - exact))
    assert e_gpu < 1e-2 and abs(e_gpu - e_cpu) < 1e-12
    ev.ctx.close()


@pytest.mark.parametrize("order", [5, 6, 7, 8])
def test_high_order_adams_bashforth(order):
    """AB5..AB8, the reference's maximum order (more old terms than the fused update carries: separate update
    kernel) incl. the self-start, vs the oracle."""
    N, dt = 4, 1e-3
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    u0 = analytic.plane_wave(x, 0.0)
    stat = np.zeros((brick.n_elements, 1, brick.n))
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u0)
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, order, 0.0, dt)
    ctx.take_steps(3)
    ev = orc.Evolution(lambda u, t: orc.dg_rhs(0, N, u, J, stat, nb), u0, 0.0, dt, f"AB{order}")
    for _ in range(3):
        ev.step()
    assert ctx.rhs_evaluations == ev.rhs_evals
    assert _relerr(ctx.get_state(), ev.u, SW_BLOCKS) < TOL
    ctx.close()


@pytest.mark.parametrize("name,stepper", [("Rk3Owren", lib.STEPPER_RK3_OWREN),
                                          ("Rk3Kennedy", lib.STEPPER_RK3_KENNEDY),
                                          ("RK4", lib.STEPPER_RK4),
                                          ("DP5", lib.STEPPER_DORMAND_PRINCE5)])
def test_butcher_tableau_steppers(name, stepper):
    """RungeKutta::update_u_impl with the reference's tableaus vs the oracle."""
    N, dt = 5, 1e-3
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    u0 = analytic.plane_wave(x, 0.0)
    stat = np.zeros((brick.n_elements, 1, brick.n))
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u0)
    ctx.set_stepper(stepper, 0, 0.0, dt)
    ctx.take_steps(3)
    ev = orc.Evolution(lambda u, t: orc.dg_rhs(0, N, u, J, stat, nb), u0, 0.0, dt, name)
    for _ in range(3):
        ev.step()
    assert ctx.rhs_evaluations == ev.rhs_evals
    assert abs(ctx.time - ev.time) < 1e-15
    assert _relerr(ctx.get_state(), ev.u, SW_BLOCKS) < TOL
    ctx.close()
