"""Edge cases and error behaviour of the C-ABI on the GPU box: the smallest and
largest supported meshes, a single element whose six faces are all external,
and the misuse errors (the reference ASSERTs / ERRORs; the library returns a
status and a message, include/dgrhs.h)."""
import ctypes

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


@pytest.mark.parametrize("N", [2, 12])
def test_single_element_all_faces_external(N):
    """One element, no neighbours: the right-hand side is the volume term only,
    and a DirichletAnalytic ghost on all six faces adds the boundary terms."""
    brick = domain.Brick([0.2, 0.1, 0.3], [0.7, 0.9, 0.8], [0, 0, 0], N,
                         periodic=(False, False, False))
    assert brick.n_elements == 1
    x, J, nbr = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    assert (nbr == -1).all()
    rng = np.random.default_rng(N)
    u = analytic.gauge_wave(x, 0.05) + 1e-3 * rng.uniform(-1, 1, (1, 50, N ** 3))
    stat = np.zeros((1, 3, N ** 3))
    stat[:, 0], stat[:, 1], stat[:, 2] = 1.0, -1.0, 1.0
    ctx = lib.Context(lib.SYSTEM_GH, N, 1, 6)
    ctx.set_geometry(J, x, nbr)
    ctx.set_static_fields(stat)
    ctx.set_state(u)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(1, N, u, J, stat, nbr)
    vol = orc.dg_rhs(1, N, u, J, stat, nbr, volume_only=True)
    np.testing.assert_array_equal(ref, vol)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    # now with ghost slots on all six faces
    nbr_g = np.array([[-(d + 2) for d in range(6)]], dtype=np.int32)
    ctx.set_geometry(J, x, nbr_g)
    f = N * N
    ghost = np.zeros((6, ctx.halo_comps, f))
    ext = np.zeros((6, 50, f))
    for d in range(6):
        p = domain._face_point_indices(N, d)
        ext[d] = analytic.gauge_wave(x[0][:, p], 0.05)
        ghost[d, :50] = ext[d]
        for i in range(3):
            ghost[d, 50 + i] = J[0][d // 2 + 3 * i, p]
        ghost[d, 53], ghost[d, 54] = stat[0][1, p], stat[0][2, p]
    ctx.set_boundary_ghost_data(0, ghost)
    ctx.compute_time_derivative(0.0)
    got = ctx.get_time_derivative()
    ref = orc.dg_rhs(1, N, u, J, stat, nbr_g, ext_u=ext)
    assert _relerr(got, ref, GH_BLOCKS) < TOL
    ctx.close()


def test_create_and_argument_errors():
    L = lib.load()
    h = ctypes.c_void_p()
    for args, msg in (((lib.SYSTEM_GH, 13, 8, 0, 0), r"n_points_1d must be in \[2, 12\]"),
                      ((lib.SYSTEM_GH, 1, 8, 0, 0), r"n_points_1d must be in \[2, 12\]"),
                      ((7, 4, 8, 0, 0), "unknown system"),
                      ((lib.SYSTEM_GH, 4, 0, 0, 0), "n_elements must be positive")):
        assert L.dgrhs_create(ctypes.byref(h), *args) != 0
        with pytest.raises(lib.DgrhsError, match=msg):
            lib._check(1)
    assert L.dgrhs_set_state(None, None) != 0
    with pytest.raises(lib.DgrhsError, match="null context"):
        lib._check(1)


def test_misuse_errors():
    N = 3
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    ctx = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, brick.n_elements)
    with pytest.raises(lib.DgrhsError, match="call dgrhs_set_geometry first"):
        ctx.set_neighbor_orientations(np.tile(np.arange(6) ^ 1, (8, 1)), np.zeros((8, 6)))
    nbr = brick.neighbors()
    bad = nbr.copy()
    bad[0, 0] = brick.n_elements
    with pytest.raises(lib.DgrhsError, match="out of range"):
        ctx.set_geometry(brick.inverse_jacobian(), None, bad)
    bad[0, 0] = -2          # a ghost slot, but the context was created without any
    with pytest.raises(lib.DgrhsError, match="ghost face index out of range"):
        ctx.set_geometry(brick.inverse_jacobian(), None, bad)
    ctx.set_geometry(brick.inverse_jacobian(), brick.coords(), nbr)
    with pytest.raises(lib.DgrhsError, match="expected 1 static components"):
        ctx.set_static_fields(np.zeros((brick.n_elements, 3, N ** 3)))
    ctx.set_static_fields(np.zeros((brick.n_elements, 1, N ** 3)))
    ctx.set_state(analytic.plane_wave(brick.coords(), 0.0))
    with pytest.raises(lib.DgrhsError, match="only apply to GH"):
        ctx.set_gauge(lib.GAUGE_DAMPED_HARMONIC, (1.0,) * 7)
    with pytest.raises(lib.DgrhsError, match="GeneralizedHarmonic boundary condition"):
        ctx.set_demand_outgoing_char_speeds(True)
    with pytest.raises(lib.DgrhsError, match="set_stepper has not been called"):
        ctx.begin_substep()
    with pytest.raises(lib.DgrhsError, match=r"order must be in \[1, 8\]"):
        ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 9, 0.0, 1e-3)
    with pytest.raises(lib.DgrhsError, match="unknown stepper"):
        ctx.set_stepper(17, 3, 0.0, 1e-3)
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 2, 0.0, 1e-3)
    with pytest.raises(lib.DgrhsError, match="end_substep without begin_substep"):
        ctx.end_substep()
    ctx.begin_substep()
    with pytest.raises(lib.DgrhsError, match="begin_substep called twice"):
        ctx.begin_substep()
    with pytest.raises(lib.DgrhsError, match="bad element range"):
        ctx.compute_time_derivative_range(0.0, 0, brick.n_elements + 1)
    with pytest.raises(lib.DgrhsError, match="dgrhs_set_interior_count"):
        ctx.compute_time_derivative_range(0.0, 0, 3)
    ctx.close()


def test_demand_outgoing_check_needs_enabling():
    N = 3
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [0, 0, 0], N, periodic=(False,) * 3)
    ctx = lib.Context(lib.SYSTEM_GH, N, 1)
    ctx.set_geometry(brick.inverse_jacobian(), brick.coords(), brick.neighbors())
    with pytest.raises(lib.DgrhsError, match="not enabled"):
        ctx.check_outgoing_char_speeds()
    # flat space, all faces external: lambda_- = -alpha < 0 everywhere
    stat = np.zeros((1, 3, N ** 3))
    stat[:, 1] = -1.0
    ctx.set_static_fields(stat)
    ctx.set_state(analytic.gauge_wave(brick.coords(), 0.0, amplitude=0.0))
    ctx.set_demand_outgoing_char_speeds(True)
    ctx.compute_time_derivative(0.0)
    n, mn = ctypes.c_longlong(0), ctypes.c_double(0.0)
    rc = lib.load().dgrhs_check_outgoing_char_speeds(ctx._h, ctypes.byref(n), ctypes.byref(mn))
    assert rc != 0 and n.value == 6 * N * N and mn.value == pytest.approx(-1.0, abs=1e-14)
    ctx.set_demand_outgoing_char_speeds(False)
    ctx.close()


def test_stream_ordered_state_transfers_double_buffered():
    """dgrhs_set_state_async / dgrhs_get_state_async: two contexts that alternate
    batches (one uploads while the other steps and downloads) return exactly what
    the blocking calls return for the same batches."""
    import torch
    N = 4
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N)
    E = brick.n_elements
    x, J, nbr = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    stat = np.zeros((E, 1, N ** 3))
    rng = np.random.default_rng(11)
    batches = [rng.uniform(-1, 1, (E, 5, N ** 3)) for _ in range(6)]

    def make():
        c = lib.Context(lib.SYSTEM_SCALAR_WAVE, N, E, 0)
        c.set_geometry(J, x, nbr)
        c.set_static_fields(stat)
        c.set_stepper(lib.STEPPER_RK3_HESTHAVEN, 3, 0.0, 1e-3)
        return c
    ref_ctx = make()
    want = []
    for b in batches:
        ref_ctx.set_state(b)
        ref_ctx.take_steps(1)
        want.append(ref_ctx.get_state())
    ref_ctx.close()
    lanes = [make(), make()]
    pinned = [[torch.empty(b.size, dtype=torch.float64, pin_memory=True) for b in batches]
              for _ in range(2)]
    outs = []
    for i, b in enumerate(batches):
        c = lanes[i % 2]
        src = pinned[0][i].numpy().reshape(b.shape)
        dst = pinned[1][i].numpy().reshape(b.shape)
        src[...] = b
        c.set_state_async(src)
        c.take_steps(1)
        c.get_state_async(dst)
        outs.append(dst)
    for c in lanes:
        c.synchronize()
    for got, ref in zip(outs, want):
        np.testing.assert_array_equal(got, ref)
    for c in lanes:
        c.close()
