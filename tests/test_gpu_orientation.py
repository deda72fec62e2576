"""Non-aligned neighbours (OrientationMap): a periodic Brick whose elements are
given independently rotated / reflected logical frames must produce exactly the
same physics.  The GPU runs the rotated mesh (with neighbour directions and
face permutations), the oracle runs the aligned mesh; results are compared after
mapping back.  Reference: orient_variables_on_slice, Domain/Structure/
OrientationMapHelpers.cpp:25-120; call site ComputeTimeDerivative.hpp:712-721."""
import itertools

import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, lib

pytestmark = pytest.mark.gpu


def _signed_perms():
    out = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            out.append((perm, signs))
    return out  # the 48 orientations of a cube


def _point_map(N, perm, signs):
    """new_index[p_old] for the frame xi'_a = signs[a] * xi_{perm[a]}."""
    p = np.arange(N ** 3)
    old = (p % N, (p // N) % N, p // (N * N))
    new = []
    for a in range(3):
        i = old[perm[a]]
        new.append(i if signs[a] > 0 else N - 1 - i)
    return new[0] + N * (new[1] + N * new[2])


def _dir_map(perm, signs):
    """old direction -> new direction."""
    m = {}
    for a in range(3):
        for side in range(2):
            old_d = 2 * perm[a] + (side if signs[a] > 0 else 1 - side)
            m[old_d] = 2 * a + side
    return m


def _face_points(N, d):
    dim, fixed = d // 2, (N - 1 if d % 2 else 0)
    q = np.arange(N * N)
    a, b = q % N, q // N
    return [fixed + N * (a + N * b), a + N * (fixed + N * b), a + N * (b + N * fixed)][dim]


def _rotate_problem(N, u, J, stat, nbr, frames):
    ne = u.shape[0]
    pm = [_point_map(N, *frames[e]) for e in range(ne)]
    dm = [_dir_map(*frames[e]) for e in range(ne)]
    u_r, J_r, s_r = np.empty_like(u), np.empty_like(J), np.empty_like(stat)
    nbr_r = np.full_like(nbr, -1)
    nd_r = np.zeros_like(nbr)
    perm_r = np.zeros_like(nbr)
    for e in range(ne):
        perm, signs = frames[e]
        u_r[e][:, pm[e]] = u[e]
        s_r[e][:, pm[e]] = stat[e]
        for a in range(3):
            for i in range(3):
                J_r[e][a + 3 * i][pm[e]] = signs[a] * J[e][perm[a] + 3 * i]
    inv_pm = [np.argsort(m) for m in pm]  # new index -> old index
    for e in range(ne):
        for d_old in range(6):
            nb = nbr[e, d_old]
            d_new = dm[e][d_old]
            nd_new = dm[nb][d_old ^ 1]
            nbr_r[e, d_new] = nb
            nd_r[e, d_new] = nd_new
            # match face points: our new face ordering -> old volume index -> the
            # aligned neighbour point (same tangential indices on the opposite face)
            fp_new = _face_points(N, d_new)
            p_old = inv_pm[e][fp_new]
            i = [p_old % N, (p_old // N) % N, p_old // (N * N)]
            dim = d_old // 2
            i[dim] = np.where(i[dim] == 0, N - 1, 0)
            p_nb_old = i[0] + N * (i[1] + N * i[2])
            p_nb_new = pm[nb][p_nb_old]
            fp_nb = _face_points(N, nd_new)
            pos = {int(v): k for k, v in enumerate(fp_nb)}
            target = np.array([pos[int(v)] for v in p_nb_new])
            q = np.arange(N * N)
            qa, qb = q % N, q // N
            found = None
            for code in range(8):
                na = np.where(code & 1, qb, qa)
                nbb = np.where(code & 1, qa, qb)
                if code & 2:
                    na = N - 1 - na
                if code & 4:
                    nbb = N - 1 - nbb
                if np.array_equal(na + N * nbb, target):
                    found = code
                    break
            assert found is not None
            perm_r[e, d_new] = found
    return u_r, J_r, s_r, nbr_r, nd_r, perm_r, pm


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
