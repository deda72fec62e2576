"""Local time stepping on the GPU (dgrhs_lts_*) against the oracle's LtsEvolution, 1e-12:
ScalarWave and GH on periodic bricks with two and three step-size levels, the Kerr-Schild
shell with non-aligned wedges and ghost boundaries, and equal levels against the GTS path."""
import numpy as np
import pytest

from oracle import lts as olts
from oracle import oracle as orc
from spectre_b200 import analytic, domain, evolution, lib
from spectre_b200 import lts as hlts
from tests.test_gpu_parity import GH_BLOCKS, SW_BLOCKS, TOL, _curved_jacobian, _relerr

pytestmark = pytest.mark.gpu


def _brick_levels(brick, x, rule):
    levels = np.array([rule(x[e].mean(axis=1)) for e in range(brick.n_elements)])
    perm, nb = hlts.order_by_level(levels, brick.neighbors())
    return levels[perm], perm, nb


def _run(system, N, J, stat, nb, levels, order, dt, u0, past, n_coarse, blocks, ostat=None,
         gauge=None, ext_u=None, nbr_dir=None, face_perm=None, ctx=None, mode=1):
    """mode 0: every internal face through the boundary histories (the reference's
    formulation term by term); mode 1 (the library's default): same-level faces like GTS
    faces, stepper update fused.  The oracle is the reference's formulation in both cases."""
    own = ctx is None
    if own:
        ctx = lib.Context(system, N, len(levels))
        ctx.set_geometry(J, None, nb)
        if nbr_dir is not None:
            ctx.set_neighbor_orientations(nbr_dir, face_perm)
        ctx.set_static_fields(stat)
        if gauge is not None:
            ctx.set_gauge(lib.GAUGE_FIELDS)
            ctx.set_gauge_fields(*gauge)
    ctx.set_state(u0)
    ctx.lts_init(order, 0.0, dt, levels, same_level_faces_in_volume_history=bool(mode))
    for j in range(1, order):
        ctx.lts_set_past_state(j, past(j))
    np.testing.assert_array_equal(ctx.get_state(), u0)
    ev = olts.LtsEvolution(system, N, J, stat if ostat is None else ostat, nb, levels, order,
                           0.0, dt, u0, past,
                           gauge_params=orc.GAUGE_HARMONIC if gauge is None else orc.GAUGE_GIVEN,
                           ext_u=ext_u, nbr_dir=nbr_dir, face_perm=face_perm)
    for _ in range(n_coarse):
        ctx.lts_take_coarse_steps(1)
        ev.take_coarse_steps(1)
        assert _relerr(ctx.get_state(), ev.u, blocks) < TOL
    t, tick = ctx.lts_time()
    assert tick == ev.tick and t == pytest.approx(ev.time(), rel=1e-14)
    if own:
        ctx.close()
    return ev


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,order,nlevels", [(4, 3, 2), (5, 2, 3), (6, 3, 3), (3, 4, 2), (8, 3, 2),
                                             (12, 3, 2), (4, 5, 2)])
def test_scalar_wave_lts(N, order, nlevels, mode):
    rng = np.random.default_rng(40 + N)
    L = 2 * np.pi
    brick = domain.Brick([0, 0, 0], [L] * 3, [1, 1, 1], N)
    x0 = brick.coords()
    rule = (lambda c: int(c[0] > np.pi)) if nlevels == 2 else \
        (lambda c: int(c[0] > np.pi) + int(c[1] > np.pi))
    levels, perm, nb = _brick_levels(brick, x0, rule)
    assert levels.max() == nlevels - 1
    x = x0[perm]
    J = _curved_jacobian(rng, brick)[perm]
    stat = rng.uniform(0, 1, (brick.n_elements, 1, brick.n))
    dt = 4e-3
    stride = 2 ** (levels.max() - levels)
    tick = dt / 2 ** levels.max()

    def past(j):
        return np.stack([analytic.plane_wave(x[e], -j * stride[e] * tick)
                         for e in range(len(levels))])
    u0 = analytic.plane_wave(x, 0.0) + 0.05 * rng.uniform(-1, 1, (brick.n_elements, 5, brick.n))
    ev = _run(lib.SYSTEM_SCALAR_WAVE, N, J, stat, nb, levels, order, dt, u0, past, 3, SW_BLOCKS,
              mode=mode)
    assert ev.corrections_evaluated > 0


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N,order,nlevels,gauge", [(4, 3, 2, False), (5, 3, 3, True), (6, 2, 2, True),
                                                   (10, 3, 2, True), (12, 4, 2, True)])
def test_gh_lts(N, order, nlevels, gauge, mode):
    rng = np.random.default_rng(60 + N)
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    x0 = brick.coords()
    rule = (lambda c: int(c[2] > 0.5)) if nlevels == 2 else \
        (lambda c: int(c[0] > 0.5) + int(c[2] > 0.5))
    levels, perm, nb = _brick_levels(brick, x0, rule)
    x = x0[perm]
    J = _curved_jacobian(rng, brick)[perm]
    stat = rng.uniform(-1, 1, (brick.n_elements, 3, brick.n))
    dt = 4e-4
    stride = 2 ** (levels.max() - levels)
    tick = dt / 2 ** levels.max()
    noise = 1e-2 * rng.uniform(-1, 1, (brick.n_elements, 50, brick.n))

    def past(j):
        return noise + np.stack([analytic.gauge_wave(x[e], 0.1 - j * stride[e] * tick)
                                 for e in range(len(levels))])
    u0 = noise + analytic.gauge_wave(x, 0.1)
    g, ostat = None, None
    if gauge:
        H = rng.uniform(-1, 1, (brick.n_elements, 4, brick.n))
        dH = rng.uniform(-1, 1, (brick.n_elements, 16, brick.n))
        g, ostat = (H, dH), np.concatenate([stat, H, dH], axis=1)
    _run(lib.SYSTEM_GH, N, J, stat, nb, levels, order, dt, u0, past, 2, GH_BLOCKS, ostat=ostat,
         gauge=g, mode=mode)


@pytest.mark.parametrize("mode", [0, 1])
def test_gh_lts_on_kerr_schild_shell(mode):
    """Two radial layers of six non-aligned wedges with DirichletAnalytic ghosts on both
    spheres; the outer layer takes two steps per step of the inner one (the element order of
    the shell is inside-out and the levels must ascend: the parity does not care which layer
    is the fine one)."""
    from tests.test_gpu_shell import _gauge_fields
    N, order, dt = 5, 3, 2e-3
    problem = evolution.gh_kerr_schild_shell_problem((0, 1), N, order="radial")
    ev0 = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ctx, part = ev0.ctx, ev0.part
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    r = np.sqrt((x ** 2).sum(axis=1)).mean(axis=1)
    levels = (r > np.median(r)).astype(np.int32)
    assert np.all(np.diff(levels) >= 0) and levels.max() == 1
    ua = problem.u0(ids, 0.0)
    rng = np.random.default_rng(5)
    u0 = ua + 1e-3 * rng.uniform(-1, 1, ua.shape)
    H, dH = _gauge_fields(N, x, J, ua)
    ext = ev0.boundary_ghost_data(problem, 0.0)[:, :50]
    _run(lib.SYSTEM_GH, N, J, stat, part.local_neighbors, levels, order, dt, u0, lambda j: u0, 2,
         GH_BLOCKS, ostat=np.concatenate([stat, H, dH], axis=1), gauge=(H, dH), ext_u=ext,
         nbr_dir=part.local_neighbor_direction, face_perm=part.local_face_permutation, ctx=ctx,
         mode=mode)
    ctx.close()


def test_equal_levels_match_the_gts_path():
    """one level: the LTS entry points take the same steps as the GTS stepper of the library
    (history started from the same past states), up to the order of the additions"""
    N, order, dt = 6, 3, 2e-3
    rng = np.random.default_rng(9)
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    stat = np.zeros((brick.n_elements, 1, brick.n))
    levels = np.zeros(brick.n_elements, dtype=np.int32)
    u0 = analytic.plane_wave(x, 0.0)
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u0)
    ctx.lts_init(order, 0.0, dt, levels)
    for j in range(1, order):
        ctx.lts_set_past_state(j, analytic.plane_wave(x, -j * dt))
    ctx.lts_take_coarse_steps(5)
    got = ctx.get_state()
    hist = [orc.dg_rhs(0, N, analytic.plane_wave(x, -j * dt), J, stat, nb) for j in (2, 1)]
    u = u0.copy()
    c = orc._AB_CONST[3]
    for _ in range(5):
        hist.append(orc.dg_rhs(0, N, u, J, stat, nb))
        u = u + dt * (c[0] * hist[-3] + c[1] * hist[-2] + c[2] * hist[-1])
    assert _relerr(got, u, SW_BLOCKS) < TOL
    assert np.max(np.abs(got - analytic.plane_wave(x, 5 * dt))) < 1e-3
    ctx.close()


def test_lts_rejections():
    N = 4
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(brick.inverse_jacobian(), brick.coords(), brick.neighbors())
    with pytest.raises(lib.DgrhsError, match="sorted by step-size level"):
        ctx.lts_init(3, 0.0, 1e-3, np.array([1, 0, 0, 0, 0, 0, 0, 0]))
    ctx.lts_init(3, 0.0, 1e-3, np.zeros(8, dtype=np.int32))
    with pytest.raises(lib.DgrhsError, match="past state 1"):
        ctx.lts_take_ticks(1)
    ctx.close()


def test_lts_started_from_a_gts_phase():
    """no analytic past states: (order - 1) coarse steps of self-started GTS with the finest
    step provide the histories (spectre_b200.lts.start_from_gts); the oracle does the same
    with its GTS Evolution and its LtsEvolution"""
    N, order, dt = 6, 3, 8e-3
    rng = np.random.default_rng(3)
    L = 2 * np.pi
    brick = domain.Brick([0, 0, 0], [L] * 3, [1, 1, 1], N)
    levels, perm, nb = _brick_levels(brick, brick.coords(),
                                     lambda c: int(c[0] > np.pi) + int(c[1] > np.pi))
    x = brick.coords()[perm]
    J = brick.inverse_jacobian()[perm]
    stat = np.zeros((brick.n_elements, 1, brick.n))
    u0 = analytic.plane_wave(x, 0.0)
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_static_fields(stat)
    ctx.set_state(u0)
    t_start = hlts.start_from_gts(ctx, order, 0.0, dt, levels)
    assert t_start == pytest.approx(2 * dt)
    # the oracle's version of the same procedure
    tick, stride = dt / 4, 2 ** (2 - levels)
    gts = orc.Evolution(lambda w, t: orc.dg_rhs(0, N, w, J, stat, nb), u0, 0.0, tick, "AB3")
    snaps = {0: u0.copy()}
    for T in range(1, 9):
        gts.step()
        snaps[T] = gts.u.copy()
    np.testing.assert_allclose(ctx.get_state(), snaps[8], rtol=0, atol=1e-12)

    def past(j):
        return np.stack([snaps[8 - j * stride[e]][e] for e in range(len(levels))])
    ev = olts.LtsEvolution(0, N, J, stat, nb, levels, order, t_start, dt, snaps[8], past)
    ctx.lts_take_coarse_steps(3)
    ev.take_coarse_steps(3)
    got = ctx.get_state()
    assert _relerr(got, ev.u, SW_BLOCKS) < TOL
    t, _ = ctx.lts_time()
    assert t == pytest.approx(5 * dt)
    assert np.max(np.abs(got - analytic.plane_wave(x, t))) < 2e-3
    ctx.close()


def test_cpp_lts_example_matches_oracle():
    """spectre_b200/host/evolve_scalar_wave_lts.cpp: the plane wave with two step-size levels
    driven entirely from C++20 through the C-ABI (dgrhs_lts_*); its ObserveNorms-style errors
    equal those of the oracle's LtsEvolution of the same configuration."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "evolve_scalar_wave_lts")
    src = os.path.join(root, "spectre_b200", "host", "evolve_scalar_wave_lts.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-o", exe, src, "-L",
                           os.path.join(root, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(root, "spectre_b200")])
    steps = 6
    out = subprocess.run([exe, str(steps)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    got = {ln.split()[0]: float(ln.split()[1]) for ln in out.stdout.strip().splitlines()}
    N, dt = 5, 2e-3
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    levels, perm, nb = _brick_levels(brick, brick.coords(), lambda c: int(c[0] > np.pi))
    x, J = brick.coords()[perm], brick.inverse_jacobian()[perm]
    stat = np.zeros((brick.n_elements, 1, N ** 3))

    def past(j):
        return np.stack([analytic.plane_wave(x[e], -j * dt / 2 ** levels[e])
                         for e in range(len(levels))])
    ev = olts.LtsEvolution(0, N, J, stat, nb, levels, 3, 0.0, dt, analytic.plane_wave(x, 0.0), past)
    ev.take_coarse_steps(steps)
    assert got["time"] == pytest.approx(ev.time(), abs=1e-15)
    exact = analytic.plane_wave(x, ev.time())
    npts = exact.shape[0] * exact.shape[2]
    for name, (a, b) in zip(("Psi", "Pi", "Phi"), ((0, 1), (1, 2), (2, 5))):
        want = np.sqrt(np.sum((ev.u[:, a:b] - exact[:, a:b]) ** 2) / npts)
        assert got[f"Error({name})"] == pytest.approx(want, rel=1e-8)


def _refined_problem(system, N, refined, rng, rule):
    """h-refined periodic brick with its mortar table, elements sorted by the step-size level
    that `rule(element size)` gives"""
    L = 2 * np.pi if system == lib.SYSTEM_SCALAR_WAVE else 1.0
    rb = domain.RefinedBrick([0, 0, 0], [L] * 3, [1, 1, 1], N, refined)
    x, nb, mt = rb.coords(), rb.neighbors(), np.array(rb.mortars())
    size = x[:, 0].max(axis=1) - x[:, 0].min(axis=1)
    levels = np.array([rule(sz / size.max()) for sz in size])
    perm, nbp = hlts.order_by_level(levels, nb)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    mtp = mt.copy()
    mtp[:, 0], mtp[:, 2] = inv[mt[:, 0]], inv[mt[:, 2]]
    J = rb.inverse_jacobian() + 0.05 * rng.uniform(-1, 1, (rb.n_elements, 9, N ** 3))
    return rb, x[perm], J[perm], nbp, mtp.astype(np.int32), levels[perm]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("system,N,order,rule", [
    ("sw", 4, 3, "fine"), ("sw", 6, 2, "fine"), ("sw", 5, 3, "mixed"), ("gh", 4, 3, "fine"),
    ("gh", 6, 3, "mixed"), ("gh", 3, 4, "coarse")])
def test_lts_with_mortars(system, N, order, rule, mode):
    """h-refinement with local time stepping: 'fine' = the children of the refined cells take
    two steps per step of the unrefined elements (the canonical set-up), 'coarse' = the other
    way round, 'mixed' = three levels that cut through the refined cells, so that one coarse
    face has mortars inside and outside the boundary histories"""
    sysid = lib.SYSTEM_GH if system == "gh" else lib.SYSTEM_SCALAR_WAVE
    rng = np.random.default_rng(17 * N + order)
    rules = {"fine": lambda s: int(s < 0.75), "coarse": lambda s: int(s > 0.75)}
    if rule == "mixed":
        counter = iter(range(10 ** 6))
        rules["mixed"] = lambda s: (next(counter) % 2) + int(s < 0.75)
    rb, x, J, nb, mt, levels = _refined_problem(sysid, N, [(0, 0, 0), (1, 1, 0)], rng, rules[rule])
    assert len(mt) > 0 and len(set(levels.tolist())) >= 2
    nelem = len(levels)
    stride = 2 ** (levels.max() - levels)
    if system == "gh":
        dt = 4e-4
        tick = dt / 2 ** levels.max()
        noise = 1e-2 * rng.uniform(-1, 1, (nelem, 50, N ** 3))
        wave = lambda xe, t: analytic.gauge_wave(xe, 0.1 + t)
        stat = rng.uniform(-1, 1, (nelem, 3, N ** 3))
        blocks = GH_BLOCKS
    else:
        dt = 4e-3
        tick = dt / 2 ** levels.max()
        noise = 0.05 * rng.uniform(-1, 1, (nelem, 5, N ** 3))
        wave = analytic.plane_wave
        stat = rng.uniform(0, 1, (nelem, 1, N ** 3))
        blocks = SW_BLOCKS

    def past(j):
        return noise + np.stack([wave(x[e], -j * stride[e] * tick) for e in range(nelem)])
    u0 = noise + np.stack([wave(x[e], 0.0) for e in range(nelem)])
    ctx = lib.Context(sysid, N, nelem)
    ctx.set_geometry(J, None, nb)
    ctx.set_mortars(mt)
    ctx.set_static_fields(stat)
    ctx.set_state(u0)
    ctx.lts_init(order, 0.0, dt, levels, same_level_faces_in_volume_history=bool(mode))
    for j in range(1, order):
        ctx.lts_set_past_state(j, past(j))
    ev = olts.LtsEvolution(0 if system == "sw" else 1, N, J, stat, nb, levels, order, 0.0, dt, u0,
                           past, mortars=mt)
    for _ in range(2):
        ctx.lts_take_coarse_steps(1)
        ev.take_coarse_steps(1)
        assert _relerr(ctx.get_state(), ev.u, blocks) < TOL
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_gh_lts_with_exponential_filter(mode):
    """dg::Actions::Filter after every element's step (KerrSchild.yaml's filter), two levels"""
    N, order, dt = 6, 3, 4e-4
    rng = np.random.default_rng(23)
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    levels, perm, nb = _brick_levels(brick, brick.coords(), lambda c: int(c[1] > 0.5))
    x = brick.coords()[perm]
    J = _curved_jacobian(rng, brick)[perm]
    stat = rng.uniform(-1, 1, (brick.n_elements, 3, brick.n))
    noise = 1e-2 * rng.uniform(-1, 1, (brick.n_elements, 50, brick.n))
    stride = 2 ** (levels.max() - levels)

    def past(j):
        return noise + np.stack([analytic.gauge_wave(x[e], 0.1 - j * stride[e] * dt / 2)
                                 for e in range(len(levels))])
    u0 = noise + analytic.gauge_wave(x, 0.1)
    alpha, half_power = 36.0, 4      # a low half power: the filter changes every mode visibly
    ctx = lib.Context(lib.SYSTEM_GH, N, brick.n_elements)
    ctx.set_geometry(J, None, nb)
    ctx.set_static_fields(stat)
    ctx.set_exponential_filter(True, alpha, half_power)
    ctx.set_state(u0)
    ctx.lts_init(order, 0.0, dt, levels, same_level_faces_in_volume_history=bool(mode))
    for j in range(1, order):
        ctx.lts_set_past_state(j, past(j))
    F = orc.exponential_filter_matrix(N, alpha, half_power)
    ev = olts.LtsEvolution(1, N, J, stat, nb, levels, order, 0.0, dt, u0, past,
                           post_update=lambda v: orc.apply_filter(N, v, F))
    plain = olts.LtsEvolution(1, N, J, stat, nb, levels, order, 0.0, dt, u0, past)
    for _ in range(2):
        ctx.lts_take_coarse_steps(1)
        ev.take_coarse_steps(1)
        assert _relerr(ctx.get_state(), ev.u, GH_BLOCKS) < TOL
    plain.take_coarse_steps(2)
    assert _relerr(plain.u, ev.u, GH_BLOCKS) > 1e-6     # the filter matters
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_gh_lts_with_the_single_black_hole_boundary_conditions(mode):
    """the boundary conditions of EvolveGhSingleBlackHole (an LTS executable in the reference):
    DemandOutgoingCharSpeeds on the excision sphere, ConstraintPreservingBjorhus (physical) on
    the outer sphere -- external boundary conditions are part of the time derivative that
    enters the element's own history (ComputeTimeDerivative applies them); two radial layers
    with different steps"""
    from tests.test_gpu_shell import _gauge_fields
    N, order, dt = 5, 3, 2e-3
    problem = evolution.gh_kerr_schild_shell_problem(
        (0, 1), N, inner_radius=1.9, outer_radius=6.0, order="radial",
        inner_boundary="DemandOutgoingCharSpeeds", outer_boundary="ConstraintPreservingPhysical")
    ev0 = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, 1e-4)
    ctx, part = ev0.ctx, ev0.part
    assert (part.local_neighbors == lib.BJORHUS_PHYSICAL).sum() == 6 and not part.external_faces
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    r = np.sqrt((x ** 2).sum(axis=1)).mean(axis=1)
    levels = (r > np.median(r)).astype(np.int32)
    assert np.all(np.diff(levels) >= 0) and levels.max() == 1
    ua = problem.u0(ids, 0.0)
    u0 = ua + 1e-3 * np.random.default_rng(6).uniform(-1, 1, ua.shape)
    H, dH = _gauge_fields(N, x, J, ua)
    ctx.set_state(u0)
    ctx.lts_init(order, 0.0, dt, levels, same_level_faces_in_volume_history=bool(mode))
    for j in range(1, order):
        ctx.lts_set_past_state(j, u0)
    ev = olts.LtsEvolution(1, N, J, np.concatenate([stat, H, dH], axis=1), part.local_neighbors,
                           levels, order, 0.0, dt, u0, lambda j: u0,
                           gauge_params=orc.GAUGE_GIVEN, nbr_dir=part.local_neighbor_direction,
                           face_perm=part.local_face_permutation, coords=x)
    for _ in range(2):
        ctx.lts_take_coarse_steps(1)
        ev.take_coarse_steps(1)
        assert _relerr(ctx.get_state(), ev.u, GH_BLOCKS) < TOL
    ctx.check_outgoing_char_speeds()
    # without the Bjorhus condition the outer layer would evolve differently
    free = olts.LtsEvolution(1, N, J, np.concatenate([stat, H, dH], axis=1),
                             np.where(part.local_neighbors == lib.BJORHUS_PHYSICAL, -1,
                                      part.local_neighbors), levels, order, 0.0, dt, u0,
                             lambda j: u0, gauge_params=orc.GAUGE_GIVEN,
                             nbr_dir=part.local_neighbor_direction,
                             face_perm=part.local_face_permutation, coords=x)
    free.take_coarse_steps(2)
    assert _relerr(free.u, ev.u, GH_BLOCKS) > 1e-9
    ctx.close()
