"""Non-conforming (2:1 h-refined) mortars on the GPU: mortar_kernel through the
C-ABI against the oracle's restatement of InternalMortarDataImpl.hpp:230-320 /
ApplyBoundaryCorrections.hpp:797-1045 (project_to_mortar, dg_boundary_terms on
the mortar, project_from_mortar, lift, add)."""
import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, lib

pytestmark = pytest.mark.gpu

TOL = 1e-12
GH_BLOCKS = [slice(0, 10), slice(10, 20), slice(20, 50)]
SW_BLOCKS = [slice(0, 1), slice(1, 2), slice(2, 5)]


def _relerr(a, b, blocks):
    return max(np.max(np.abs(a[:, s] - b[:, s])) / np.max(np.abs(b[:, s])) for s in blocks)


def _setup(system, N, refined, periodic=True, seed=0):
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N, refined,
                             periodic=(periodic,) * 3)
    x, nb, mt = rb.coords(), rb.neighbors(), rb.mortars()
    rng = np.random.default_rng(seed)
    # a full, per-point perturbed inverse Jacobian (faces of neighbours need not
    # see the same J: each side uses its own normal, as on a curved mesh)
    J = rb.inverse_jacobian() + 0.05 * rng.uniform(-1, 1, (rb.n_elements, 9, N ** 3))
    if system == "gh":
        u = analytic.gauge_wave(x, 0.1) + 1e-2 * rng.uniform(-1, 1, (rb.n_elements, 50, N ** 3))
        stat = rng.uniform(-1, 1, (rb.n_elements, 3, N ** 3))
        sysid = lib.SYSTEM_GH
    else:
        u = analytic.plane_wave(x, 0.3) + 0.1 * rng.uniform(-1, 1, (rb.n_elements, 5, N ** 3))
        stat = rng.uniform(0, 1, (rb.n_elements, 1, N ** 3))
        sysid = lib.SYSTEM_SCALAR_WAVE
    ctx = lib.Context(sysid, N, rb.n_elements)
    ctx.set_geometry(J, x, nb)
    ctx.set_mortars(mt)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    return rb, ctx, u, J, stat, nb, mt


@pytest.mark.parametrize("system,N", [("sw", 2), ("sw", 4), ("sw", 7), ("gh", 3), ("gh", 6),
                                      ("gh", 8), ("gh", 12)])
def test_rhs_with_mortars_matches_oracle(system, N):
    refined = [(0, 0, 0), (1, 1, 0)] if N < 12 else [(1, 0, 1)]
    rb, ctx, u, J, stat, nb, mt = _setup(system, N, refined, seed=N)
    assert len(mt) > 0
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    sid = 1 if system == "gh" else 0
    ref = orc.dg_rhs(sid, N, u, J, stat, nb, mortars=mt)
    blocks = GH_BLOCKS if system == "gh" else SW_BLOCKS
    assert _relerr(got, ref, blocks) < TOL
    # the mortars matter: dropping them changes the answer
    nomortar = orc.dg_rhs(sid, N, u, J, stat, np.where(nb == domain.HANGING, -1, nb))
    assert _relerr(nomortar, ref, blocks) > 1e-4
    ctx.close()


@pytest.mark.parametrize("system,N", [("sw", 5), ("gh", 4)])
def test_anisotropic_refinement_matches_oracle(system, N):
    """Cells split in one or two dimensions: mortars with MortarSize Full in one
    face dimension (identity projection in that dimension)."""
    split = {(0, 0, 0): (True, False, False), (1, 1, 1): (True, True, False),
             (1, 0, 0): (False, False, True)}
    rb, ctx, u, J, stat, nb, mt = _setup(system, N, split, seed=20 + N)
    assert {(int(m[4]), int(m[5])) for m in mt} >= {(1, 0), (0, 1), (1, 1)}
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(1 if system == "gh" else 0, N, u, J, stat, nb, mortars=mt)
    assert _relerr(got, ref, GH_BLOCKS if system == "gh" else SW_BLOCKS) < TOL
    ctx.close()


@pytest.mark.parametrize("system", ["sw", "gh"])
def test_evolution_with_mortars(system):
    N, dt = 4, 2e-4
    rb, ctx, u, J, stat, nb, mt = _setup(system, N, [(1, 0, 0)], seed=3)
    sid = 1 if system == "gh" else 0
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
    ctx.take_steps(2)
    ev = orc.Evolution(lambda v, t: orc.dg_rhs(sid, N, v, J, stat, nb, mortars=mt), u, 0.0, dt,
                       "AB3")
    ev.step()
    ev.step()
    assert ctx.rhs_evaluations == ev.rhs_evals
    assert _relerr(ctx.get_state(), ev.u, GH_BLOCKS if system == "gh" else SW_BLOCKS) < TOL
    ctx.close()


def test_smooth_solution_sees_refinement_only_at_truncation_level():
    """A resolved plane wave on the refined mesh: the right-hand side on the
    coarse elements next to the refined cell differs from the conforming mesh's
    only by the (small) interpolation error."""
    N = 8
    rb = domain.RefinedBrick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N, [(0, 0, 0)])
    base = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    out = {}
    for name, dom in (("refined", rb), ("base", base)):
        x, J, nb = dom.coords(), dom.inverse_jacobian(), dom.neighbors()
        ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, dom.n_elements)
        ctx.set_geometry(J, x, nb)
        if name == "refined":
            ctx.set_mortars(dom.mortars())
        ctx.set_static_fields(np.zeros((dom.n_elements, 1, N ** 3)))
        ctx.set_state(analytic.plane_wave(x, 0.0))
        ctx.compute_time_derivative(0.0)
        out[name] = ctx.get_time_derivative()
        ctx.close()
    worst = 0.0
    for e, (c, ch) in enumerate(rb.elements):
        if ch is None:
            worst = max(worst, np.max(np.abs(out["refined"][e] - out["base"][base.index_of[c]])))
    # 4e-4 for 8 points per half wavelength-and-a-half; the right-hand side itself is O(1)
    assert 0.0 < worst < 2e-3, worst


def test_mortar_table_validation():
    N = 3
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N, [(0, 0, 0)])
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, rb.n_elements)
    with pytest.raises(lib.DgrhsError, match="call dgrhs_set_geometry first"):
        ctx.set_mortars(rb.mortars())
    ctx.set_geometry(rb.inverse_jacobian(), rb.coords(), rb.neighbors())
    mt = rb.mortars().copy()
    with pytest.raises(lib.DgrhsError, match="marked hanging but the mortar table covers"):
        ctx.set_mortars(mt[:-1])
    bad = mt.copy()
    bad[0, 3] = bad[0, 1]   # that face of the fine element is not a hanging one
    with pytest.raises(lib.DgrhsError, match="must be marked DGRHS_NEIGHBOR_HANGING"):
        ctx.set_mortars(bad)
    bad = mt.copy()
    bad[0, 3] |= 8 << 3
    with pytest.raises(lib.DgrhsError, match="face permutation out of range"):
        ctx.set_mortars(bad)
    bad = mt.copy()
    bad[0, 4] = 3
    with pytest.raises(lib.DgrhsError, match="bad mortar size"):
        ctx.set_mortars(bad)
    bad = mt.copy()
    bad[1, 2] = bad[0, 2]
    with pytest.raises(lib.DgrhsError, match="listed twice"):
        ctx.set_mortars(bad)
    ctx.set_mortars(mt)
    ctx.close()


def _emulate_ranks(problem, world, steps, dt):
    """`world` contexts on one GPU, halo moved by device copies (as in
    test_gpu_parity.test_partitioned_evolution_matches_single_context)."""
    import torch
    from spectre_b200 import evolution
    evs = [evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 2, dt, device=0,
                               world=world, rank=r) for r in range(world)]
    N = problem.N
    per_face = evs[0].ctx.halo_comps * N * N

    def offsets(counts):
        out, o = [], 0
        for cnt in counts:
            out.append(o)
            o += cnt
        return out
    done = 0
    while done < steps:
        times = [ev.ctx.begin_substep() for ev in evs]
        for ev in evs:
            ev.ctx.pack_halo()
            ev.ctx.synchronize()
        for r in range(world):
            ro = offsets(evs[r].part.recv_counts)
            for p in range(world):
                cnt = evs[r].part.recv_counts[p]
                if cnt:
                    so = offsets(evs[p].part.send_counts)[r]
                    evs[r]._recv[ro[p] * per_face:(ro[p] + cnt) * per_face].copy_(
                        evs[p]._send[so * per_face:(so + cnt) * per_face])
        torch.cuda.synchronize()
        fin = []
        for ev in evs:
            if ev.part.n_interior > 0:
                ev.ctx.compute_time_derivative_range(times[0], 0, ev.part.n_interior)
            ev.ctx.compute_time_derivative_range(times[0], ev.part.n_interior, ev.part.n_local)
            fin.append(ev.ctx.end_substep())
        done += int(fin[0])
    out = None
    for ev in evs:
        st = ev.ctx.get_state()
        if out is None:
            out = np.empty((problem.brick.n_elements,) + st.shape[1:])
        out[ev.part.global_ids] = st
        ev.ctx.close()
    return out


@pytest.mark.parametrize("kind,world", [("brick", 2), ("brick", 3), ("shell", 2),
                                        ("wedge-refined shell", 2)])
def test_mortars_across_ranks_match_single_context(kind, world):
    """Mortars whose two sides live on different ranks: the remote side's face
    arrives in a ghost slot, the coarse side's rank projects and sums, the fine
    side's rank lifts its own face -- bit-identical to the single-context run."""
    from spectre_b200 import evolution
    if kind == "brick":
        N, dt = 4, 2e-4
        rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N,
                                 {(0, 0, 0): (True, True, True), (1, 1, 0): (True, False, True)})
        problem = evolution.Problem(lib.SYSTEM_GH, rb, lambda x, t: analytic.gauge_wave(x, t),
                                    (1.0, -1.0, 1.0))
    elif kind == "shell":
        N, dt = 3, 1e-4
        problem = evolution.gh_kerr_schild_shell_problem([(2, 0), (1, 1)], N,
                                                         radial_partitioning=(2.1,))
    else:   # mortars between blocks that are not aligned (rows with face permutations)
        N, dt = 3, 1e-4
        problem = evolution.gh_kerr_schild_shell_problem(
            [[(1, 1), (0, 0), (0, 0), (0, 0), (1, 0), (0, 0)]], N)
        assert ((np.asarray(problem.mortars)[:, 3] >> 3) != 0).any()
    assert len(problem.mortars) > 0
    single = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 2, dt)
    single.take_steps(2)
    ref = single.gather_state(problem.brick.n_elements)
    single.ctx.close()
    parts = [domain.Partition(problem.neighbors, world, r, mortars=problem.mortars)
             for r in range(world)]
    assert any((p.local_mortars[:, [0, 2]] < 0).any() for p in parts)   # really across ranks
    got = _emulate_ranks(problem, world, 2, dt)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("system,N", [("sw", 4), ("gh", 4), ("gh", 7)])
def test_mortars_between_non_aligned_blocks(system, N):
    """Every element of an h-refined Brick in its own rotated / reflected logical frame
    (tests/rotation.py): oriented conforming faces, oriented mortar rows and mortar
    sizes in the rotated coarse frames.  The GPU runs the rotated mesh, the oracle the
    aligned one (tests/test_mortars_cpu.py shows the oracle agrees with itself)."""
    from tests import rotation
    rng = np.random.default_rng(N)
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N,
                             {(0, 0, 0): (1, 1, 1), (1, 1, 0): (1, 0, 1), (0, 1, 1): (0, 0, 1)})
    x, nb, mt = rb.coords(), rb.neighbors(), rb.mortars()
    E = rb.n_elements
    J = rb.inverse_jacobian() + 0.05 * rng.uniform(-1, 1, (E, 9, N ** 3))
    if system == "gh":
        sysid, sid, blocks = lib.SYSTEM_GH, 1, GH_BLOCKS
        u = analytic.gauge_wave(x, 0.1) + 1e-2 * rng.uniform(-1, 1, (E, 50, N ** 3))
        stat = rng.uniform(-1, 1, (E, 3, N ** 3))
    else:
        sysid, sid, blocks = lib.SYSTEM_SCALAR_WAVE, 0, SW_BLOCKS
        u = analytic.plane_wave(x, 0.3) + 0.1 * rng.uniform(-1, 1, (E, 5, N ** 3))
        stat = rng.uniform(0, 1, (E, 1, N ** 3))
    all48 = rotation.signed_perms()
    frames = [all48[k] for k in rng.choice(48, E)]
    u_r, J_r, s_r, nbr_r, nd_r, perm_r, pm, mt_r = rotation.rotate_problem(
        N, u, J, stat, nb, frames, mortars=mt)
    assert ((mt_r[:, 3] >> 3) != 0).any()
    ctx = lib.Context(sysid, N, E)
    ctx.set_geometry(J_r, None, nbr_r)
    ctx.set_neighbor_orientations(nd_r, perm_r)
    ctx.set_mortars(mt_r)
    ctx.set_static_fields(s_r)
    ctx.set_state(u_r)
    ctx.compute_time_derivative(0.0)
    got_r = ctx.get_time_derivative()
    got = np.empty_like(got_r)
    for e in range(E):
        got[e] = got_r[e][:, pm[e]]
    ref = orc.dg_rhs(sid, N, u, J, stat, nb, mortars=mt)
    assert _relerr(got, ref, blocks) < TOL
    ref_r = orc.dg_rhs(sid, N, u_r, J_r, s_r, nbr_r, nbr_dir=nd_r, face_perm=perm_r, mortars=mt_r)
    assert _relerr(got_r, ref_r, blocks) < TOL
    ctx.close()
