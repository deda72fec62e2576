"""ctypes binding of the C-ABI in include/dgrhs.h (libdgrhs.so, built in-tree by
``__graft_entry__.build()``).  There is no CPU fallback: importing works without
a GPU (so the symbol table can be checked), but creating a context without a
CUDA device raises."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DGRHS_LIB lets a developer point at an experimental build of the same C-ABI
LIB_PATH = os.environ.get("DGRHS_LIB") or os.path.join(_HERE, "libdgrhs.so")

SYSTEM_SCALAR_WAVE, SYSTEM_GH = 0, 1
GAUGE_HARMONIC, GAUGE_FIELDS, GAUGE_DAMPED_HARMONIC, GAUGE_ANALYTIC_GAUGE_WAVE = 0, 1, 2, 3
STEPPER_ADAMS_BASHFORTH, STEPPER_RK3_HESTHAVEN = 0, 1
STEPPER_RK3_OWREN, STEPPER_RK3_KENNEDY, STEPPER_RK4, STEPPER_DORMAND_PRINCE5 = 2, 3, 4, 5

# every symbol include/dgrhs.h declares
EXPORTS = [
    "dgrhs_last_error", "dgrhs_kernel_launch_count", "dgrhs_create", "dgrhs_destroy",
    "dgrhs_set_geometry", "dgrhs_set_neighbor_orientations", "dgrhs_set_static_fields", "dgrhs_set_gauge",
    "dgrhs_set_gauge_fields", "dgrhs_set_gauge_analytic_christoffel",
    "dgrhs_set_boundary_ghost_data", "dgrhs_set_state", "dgrhs_get_state",
    "dgrhs_set_state_async", "dgrhs_get_state_async", "dgrhs_stepper_properties",
    "dgrhs_projection_matrix_meshes",
    "dgrhs_get_time_derivative", "dgrhs_compute_time_derivative", "dgrhs_set_interior_count",
    "dgrhs_pack_halo", "dgrhs_compute_time_derivative_range", "dgrhs_set_halo_map",
    "dgrhs_halo_send_ptr", "dgrhs_halo_recv_ptr", "dgrhs_halo_comps", "dgrhs_set_stepper",
    "dgrhs_take_steps", "dgrhs_time", "dgrhs_rhs_evaluations", "dgrhs_begin_substep",
    "dgrhs_end_substep", "dgrhs_set_exponential_filter", "dgrhs_exponential_filter_matrix",
    "dgrhs_set_demand_outgoing_char_speeds", "dgrhs_check_outgoing_char_speeds",
    "dgrhs_set_mortars", "dgrhs_projection_matrix",
    "dgrhs_set_fused_update", "dgrhs_set_split_volume", "dgrhs_time_kernels", "dgrhs_gh_constraint_norms",
    "dgrhs_synchronize", "dgrhs_stream", "dgrhs_state_device_ptr",
    "dgrhs_padded_points", "dgrhs_partial_derivatives", "dgrhs_differentiation_matrix",
    "dgrhs_collocation_points_and_weights", "dgrhs_adams_bashforth_coefficients",
    "dgrhs_gh_time_derivative", "dgrhs_gh_bjorhus_dg_time_derivative", "dgrhs_sw_time_derivative", "dgrhs_gh_package_data",
    "dgrhs_gh_boundary_terms", "dgrhs_sw_package_data", "dgrhs_sw_boundary_terms",
    "dgrhs_lift_flux",
    "dgrhs_comm_unique_id", "dgrhs_comm_init", "dgrhs_set_halo_peers", "dgrhs_exchange_halo",
    "dgrhs_set_phase_timing", "dgrhs_get_phase_times",
    "dgrhs_apply_exponential_filter", "dgrhs_butcher_row", "dgrhs_update_u",
    "dgrhs_project_to_mortar", "dgrhs_project_from_mortar", "dgrhs_orient_variables_on_slice",
    "dgrhs_set_p_mortars", "dgrhs_p_mortar_transfer", "dgrhs_set_slab", "dgrhs_self_start_substeps_left", "dgrhs_stepper_substep_fractions",
    "dgrhs_set_mesh_velocity", "dgrhs_gh_package_data_moving", "dgrhs_sw_package_data_moving",
    "dgrhs_adams_lts_coefficients", "dgrhs_lts_init", "dgrhs_lts_set_past_state",
    "dgrhs_lts_take_ticks", "dgrhs_lts_ticks_per_coarse_step", "dgrhs_lts_time",
    "dgrhs_lts_set_mode", "dgrhs_adams_lts_coefficients_general",
]

_lib = None


HANGING = -2 ** 31   # DGRHS_NEIGHBOR_HANGING
BJORHUS = -2 ** 31 + 1   # DGRHS_NEIGHBOR_BJORHUS (Type ConstraintPreserving)
BJORHUS_PHYSICAL = -2 ** 31 + 2   # DGRHS_NEIGHBOR_BJORHUS_PHYSICAL
P_MORTAR = -2 ** 31 + 3   # DGRHS_NEIGHBOR_P_MORTAR (neighbour with a different N)


def stepper_properties(stepper, order=0):
    """(order, number_of_substeps, number_of_past_steps, stable_step) of a time stepper
    (TimeStepper.hpp:47-246); host-only."""
    o, s, p = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    st = ctypes.c_double()
    _check(load().dgrhs_stepper_properties(int(stepper), int(order), ctypes.byref(o),
                                           ctypes.byref(s), ctypes.byref(p), ctypes.byref(st)))
    return o.value, s.value, p.value, st.value


def projection_matrix(N, child_to_parent, size):
    """Spectral::projection_matrix_{parent_to_child,child_to_parent}; size 0 Full,
    1 LowerHalf, 2 UpperHalf (host function, no GPU needed)."""
    M = np.zeros((N, N))
    _check(load().dgrhs_projection_matrix(N, int(child_to_parent), size, _ptr(M)))
    return M


def projection_matrix_meshes(n_parent, n_child, child_to_parent, size):
    """The same between meshes with different numbers of points (the child / mortar
    mesh is the finer one): [n_child, n_parent] or, child to parent, [n_parent, n_child]."""
    M = np.zeros((n_parent, n_child) if child_to_parent else (n_child, n_parent))
    _check(load().dgrhs_projection_matrix_meshes(n_parent, n_child, int(child_to_parent), size,
                                                 _ptr(M)))
    return M


class DgrhsError(RuntimeError):
    pass


def adams_lts_coefficients(local_ticks, remote_ticks, start, end, local_order, remote_order=None,
                           small_order=None, origin=0.0, tick_size=1.0, local_implicit=False,
                           remote_implicit=False, small_implicit=False):
    """adams_lts::lts_coefficients (AdamsLts.cpp:330-437), host only: {(local id, remote id):
    coefficient}.  An id is a tick (step id) or a pair (tick, step size): the substep
    (predictor) id of that step, which the implicit (Adams-Moulton) schemes use."""
    remote_order = local_order if remote_order is None else remote_order
    small_order = local_order if small_order is None else small_order

    def split(ids):
        t = np.ascontiguousarray([i[0] if isinstance(i, tuple) else i for i in ids], dtype=np.int64)
        s = np.ascontiguousarray([i[1] if isinstance(i, tuple) else 0 for i in ids], dtype=np.int64)
        return t, s
    lt, ls = split(local_ticks)
    rt, rs = split(remote_ticks)
    cap = 256
    li, ri = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
    cf = np.zeros(cap)
    n = ctypes.c_int()
    _check(load().dgrhs_adams_lts_coefficients_general(
        int(local_implicit), int(local_order), int(remote_implicit), int(remote_order),
        int(small_implicit), int(small_order), len(lt), _ptr(lt), _ptr(ls), len(rt), _ptr(rt),
        _ptr(rs), ctypes.c_longlong(int(start)), ctypes.c_longlong(int(end)),
        ctypes.c_double(origin), ctypes.c_double(tick_size), cap, ctypes.byref(n), _ptr(li),
        _ptr(ri), _ptr(cf)))
    key = lambda t, s, i: int(t[i]) if s[i] == 0 else (int(t[i]), int(s[i]))
    return {(key(lt, ls, li[t]), key(rt, rs, ri[t])): float(cf[t]) for t in range(n.value)}


def load():
    """Load libdgrhs.so; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DgrhsError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.dgrhs_last_error.restype = ctypes.c_char_p
        lib.dgrhs_kernel_launch_count.restype = ctypes.c_int64
        lib.dgrhs_rhs_evaluations.restype = ctypes.c_int64
        lib.dgrhs_time.restype = ctypes.c_double
        for f in ("dgrhs_halo_send_ptr", "dgrhs_halo_recv_ptr", "dgrhs_stream",
                  "dgrhs_state_device_ptr"):
            getattr(lib, f).restype = ctypes.c_void_p
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise DgrhsError(load().dgrhs_last_error().decode())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def comm_unique_id() -> bytes:
    """ncclUniqueId (128 bytes) for dgrhs_comm_init: rank 0 makes it, every rank gets it."""
    buf = ctypes.create_string_buffer(128)
    _check(load().dgrhs_comm_unique_id(buf))
    return buf.raw


def kernel_launch_count() -> int:
    return int(load().dgrhs_kernel_launch_count())


def differentiation_matrix(N: int) -> np.ndarray:
    D = np.zeros((N, N))
    _check(load().dgrhs_differentiation_matrix(N, _ptr(D)))
    return D


def collocation_points_and_weights(N: int):
    x, w = np.zeros(N), np.zeros(N)
    _check(load().dgrhs_collocation_points_and_weights(N, _ptr(x), _ptr(w)))
    return x, w


def exponential_filter_matrix(N: int, alpha: float, half_power: int) -> np.ndarray:
    F = np.zeros((N, N))
    _check(load().dgrhs_exponential_filter_matrix(N, ctypes.c_double(alpha), half_power, _ptr(F)))
    return F


def adams_bashforth_coefficients(times, step_start, step_end):
    t = _f64(times)
    c = np.zeros(len(t))
    _check(load().dgrhs_adams_bashforth_coefficients(
        len(t), _ptr(t), ctypes.c_double(step_start), ctypes.c_double(step_end), _ptr(c)))
    return c


def partial_derivatives(N: int, u: np.ndarray, inv_jacobian: np.ndarray) -> np.ndarray:
    u, J = _f64(u), _f64(inv_jacobian)
    C = u.shape[0]
    du = np.zeros((3 * C, N ** 3))
    _check(load().dgrhs_partial_derivatives(N, C, _ptr(u), _ptr(J), _ptr(du)))
    return du


def gh_time_derivative(u, du, gamma0, gamma1, gamma2, gauge_h=None, d4_gauge_h=None):
    u, du = _f64(u), _f64(du)
    n = u.shape[1]
    g = [_f64(x) for x in (gamma0, gamma1, gamma2)]
    harmonic = gauge_h is None
    H = None if harmonic else _f64(gauge_h)
    dH = None if harmonic else _f64(d4_gauge_h)
    dt = np.zeros((50, n))
    _check(load().dgrhs_gh_time_derivative(n, _ptr(u), _ptr(du), _ptr(g[0]), _ptr(g[1]),
                                           _ptr(g[2]), int(harmonic), _ptr(H), _ptr(dH),
                                           _ptr(dt)))
    return dt


def sw_time_derivative(u, du, gamma2):
    u, du, gamma2 = _f64(u), _f64(du), _f64(gamma2)
    n = u.shape[1]
    dt = np.zeros((5, n))
    _check(load().dgrhs_sw_time_derivative(n, _ptr(u), _ptr(du), _ptr(gamma2), _ptr(dt)))
    return dt


def gh_package_data(u, gamma1, gamma2, lapse, shift, normal_covector, normal_vector,
                    normal_dot_mesh_velocity=None):
    a = [_f64(x) for x in (u, gamma1, gamma2, lapse, shift, normal_covector, normal_vector)]
    f = a[0].shape[1]
    pk = np.zeros((134, f))
    ms = ctypes.c_double()
    nv = None if normal_dot_mesh_velocity is None else _f64(normal_dot_mesh_velocity)
    _check(load().dgrhs_gh_package_data_moving(f, *[_ptr(x) for x in a], _ptr(nv), _ptr(pk),
                                               ctypes.byref(ms)))
    return pk, ms.value


def gh_boundary_terms(pk_int, pk_ext):
    a, b = _f64(pk_int), _f64(pk_ext)
    f = a.shape[1]
    c = np.zeros((50, f))
    _check(load().dgrhs_gh_boundary_terms(f, _ptr(a), _ptr(b), _ptr(c)))
    return c


def sw_package_data(u, gamma2, normal_covector, normal_dot_mesh_velocity=None):
    a = [_f64(x) for x in (u, gamma2, normal_covector)]
    f = a[0].shape[1]
    pk = np.zeros((16, f))
    ms = ctypes.c_double()
    nv = None if normal_dot_mesh_velocity is None else _f64(normal_dot_mesh_velocity)
    _check(load().dgrhs_sw_package_data_moving(f, *[_ptr(x) for x in a], _ptr(nv), _ptr(pk),
                                               ctypes.byref(ms)))
    return pk, ms.value


def sw_boundary_terms(pk_int, pk_ext):
    a, b = _f64(pk_int), _f64(pk_ext)
    f = a.shape[1]
    c = np.zeros((5, f))
    _check(load().dgrhs_sw_boundary_terms(f, _ptr(a), _ptr(b), _ptr(c)))
    return c


def lift_flux(corr, extent, magnitude):
    c, m = _f64(corr).copy(), _f64(magnitude)
    _check(load().dgrhs_lift_flux(c.shape[1], c.shape[0], _ptr(c), extent, _ptr(m)))
    return c


class Context:
    """One batched DG context per GPU (dgrhs_ctx)."""

    def __init__(self, system: int, N: int, n_elements: int, n_ghost_faces: int = 0,
                 device: int = 0):
        self._lib = load()
        self._h = ctypes.c_void_p()
        self.system, self.N, self.n_elements = system, N, n_elements
        self.n_ghost_faces = n_ghost_faces
        self.n_vars = 50 if system == SYSTEM_GH else 5
        self.n = N ** 3
        _check(self._lib.dgrhs_create(ctypes.byref(self._h), system, N, n_elements,
                                      n_ghost_faces, device))

    def close(self):
        if self._h:
            self._lib.dgrhs_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_geometry(self, inv_jacobian, coords, neighbors):
        J = _f64(inv_jacobian)
        X = _f64(coords) if coords is not None else None
        nb = np.ascontiguousarray(neighbors, dtype=np.int32)
        assert J.shape == (self.n_elements, 9, self.n)
        assert nb.shape == (self.n_elements, 6)
        _check(self._lib.dgrhs_set_geometry(self._h, _ptr(J), _ptr(X), _ptr(nb)))

    def set_neighbor_orientations(self, neighbor_direction, face_permutation):
        nd = np.ascontiguousarray(neighbor_direction, dtype=np.int32)
        pm = np.ascontiguousarray(face_permutation, dtype=np.int32)
        assert nd.shape == pm.shape == (self.n_elements, 6)
        _check(self._lib.dgrhs_set_neighbor_orientations(self._h, _ptr(nd), _ptr(pm)))

    def set_static_fields(self, fields):
        F = _f64(fields)
        assert F.shape[0] == self.n_elements and F.shape[2] == self.n
        _check(self._lib.dgrhs_set_static_fields(self._h, _ptr(F), F.shape[1]))

    def lts_init(self, order, t0, dt_coarse, levels, same_level_faces_in_volume_history=True):
        """Adams-Bashforth local time stepping with steps dt_coarse / 2^levels[e] (levels
        ascending in the element order)."""
        _check(self._lib.dgrhs_lts_set_mode(self._h, int(same_level_faces_in_volume_history)))
        lv = np.ascontiguousarray(levels, dtype=np.int32)
        assert lv.shape == (self.n_elements,)
        _check(self._lib.dgrhs_lts_init(self._h, int(order), ctypes.c_double(t0),
                                        ctypes.c_double(dt_coarse), _ptr(lv)))

    def lts_set_past_state(self, j, u):
        U = _f64(u)
        assert U.shape == (self.n_elements, self.n_vars, self.n)
        _check(self._lib.dgrhs_lts_set_past_state(self._h, int(j), _ptr(U)))

    def lts_take_ticks(self, n):
        _check(self._lib.dgrhs_lts_take_ticks(self._h, ctypes.c_longlong(int(n))))

    def lts_take_coarse_steps(self, n):
        k = ctypes.c_longlong()
        _check(self._lib.dgrhs_lts_ticks_per_coarse_step(self._h, ctypes.byref(k)))
        self.lts_take_ticks(n * k.value)

    def lts_time(self):
        t, k = ctypes.c_double(), ctypes.c_longlong()
        _check(self._lib.dgrhs_lts_time(self._h, ctypes.byref(t), ctypes.byref(k)))
        return t.value, k.value

    def set_mesh_velocity(self, v):
        """Inertial mesh velocity [n_elements, 3, n] of a moving mesh, or None (static)."""
        if v is None:
            _check(self._lib.dgrhs_set_mesh_velocity(self._h, None))
            return
        V = _f64(v)
        assert V.shape == (self.n_elements, 3, self.n)
        _check(self._lib.dgrhs_set_mesh_velocity(self._h, _ptr(V)))

    def set_gauge(self, gauge, params=()):
        p = _f64(list(params)) if len(params) else np.zeros(1)
        _check(self._lib.dgrhs_set_gauge(self._h, gauge, _ptr(p), len(params)))

    def set_gauge_fields(self, H, dH):
        H, dH = _f64(H), _f64(dH)
        assert H.shape == (self.n_elements, 4, self.n)
        assert dH.shape == (self.n_elements, 16, self.n)
        _check(self._lib.dgrhs_set_gauge_fields(self._h, _ptr(H), _ptr(dH)))

    def set_state(self, u):
        u = _f64(u)
        assert u.shape == (self.n_elements, self.n_vars, self.n), u.shape
        _check(self._lib.dgrhs_set_state(self._h, _ptr(u)))

    def get_state(self):
        u = np.zeros((self.n_elements, self.n_vars, self.n))
        _check(self._lib.dgrhs_get_state(self._h, _ptr(u)))
        return u

    def set_state_async(self, u):
        """Stream-ordered upload; `u` ([E][vars][n] float64, C-contiguous, ideally
        page-locked) must stay alive until `synchronize()`."""
        assert u.dtype == np.float64 and u.flags.c_contiguous
        assert u.shape == (self.n_elements, self.n_vars, self.n), u.shape
        _check(self._lib.dgrhs_set_state_async(self._h, _ptr(u)))

    def get_state_async(self, out):
        assert out.dtype == np.float64 and out.flags.c_contiguous
        assert out.shape == (self.n_elements, self.n_vars, self.n), out.shape
        _check(self._lib.dgrhs_get_state_async(self._h, _ptr(out)))

    def get_time_derivative(self):
        u = np.zeros((self.n_elements, self.n_vars, self.n))
        _check(self._lib.dgrhs_get_time_derivative(self._h, _ptr(u)))
        return u

    def compute_time_derivative(self, time=0.0, volume_only=False):
        _check(self._lib.dgrhs_compute_time_derivative(self._h, ctypes.c_double(time),
                                                       int(volume_only)))

    def compute_time_derivative_range(self, time, begin, end):
        _check(self._lib.dgrhs_compute_time_derivative_range(self._h, ctypes.c_double(time),
                                                             begin, end))

    def set_interior_count(self, n_interior):
        _check(self._lib.dgrhs_set_interior_count(self._h, int(n_interior)))

    def set_halo_map(self, ghost_send_map):
        m = np.ascontiguousarray(ghost_send_map, dtype=np.int32).reshape(-1, 2)
        _check(self._lib.dgrhs_set_halo_map(self._h, _ptr(m) if len(m) else None, len(m)))

    def set_boundary_ghost_data(self, slot_begin, data):
        d = _f64(data)
        assert d.shape[1:] == (self.halo_comps, self.N ** 2), d.shape
        _check(self._lib.dgrhs_set_boundary_ghost_data(self._h, slot_begin, d.shape[0], _ptr(d)))

    def set_gauge_analytic_christoffel(self, u_analytic):
        u = _f64(u_analytic)
        assert u.shape == (self.n_elements, 50, self.n)
        _check(self._lib.dgrhs_set_gauge_analytic_christoffel(self._h, _ptr(u)))

    def pack_halo(self):
        _check(self._lib.dgrhs_pack_halo(self._h))

    @property
    def halo_comps(self):
        return int(self._lib.dgrhs_halo_comps(self._h))

    def halo_send_ptr(self):
        return self._lib.dgrhs_halo_send_ptr(self._h)

    def halo_recv_ptr(self):
        return self._lib.dgrhs_halo_recv_ptr(self._h)

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Join the NCCL communicator of the ranks that share the domain (collective)."""
        assert len(unique_id) == 128
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        _check(self._lib.dgrhs_comm_init(self._h, buf, int(rank), int(world)))

    def set_halo_peers(self, send_counts, recv_counts):
        sc = np.ascontiguousarray(send_counts, dtype=np.int32)
        rc = np.ascontiguousarray(recv_counts, dtype=np.int32)
        _check(self._lib.dgrhs_set_halo_peers(self._h, sc.ctypes.data_as(ctypes.c_void_p),
                                              rc.ctypes.data_as(ctypes.c_void_p)))

    def set_phase_timing(self, enable=True):
        _check(self._lib.dgrhs_set_phase_timing(self._h, int(bool(enable))))

    def phase_times(self):
        """ms from the start of the last multi-GPU RHS evaluation to the end of (pack,
        interior faces, interior volume, NCCL start, NCCL end, remaining faces, boundary
        volume)."""
        ms = np.zeros(7)
        _check(self._lib.dgrhs_get_phase_times(self._h, _ptr(ms)))
        return dict(zip(("pack", "faces_interior", "volume_interior", "nccl_start", "nccl_end",
                         "faces_boundary", "volume_boundary"), ms.tolist()))

    def exchange_halo(self):
        _check(self._lib.dgrhs_exchange_halo(self._h))

    def set_stepper(self, stepper, order, t0, dt):
        _check(self._lib.dgrhs_set_stepper(self._h, stepper, order, ctypes.c_double(t0),
                                           ctypes.c_double(dt)))

    def set_slab(self, slab_start, slab_end, steps_per_slab=1):
        """Exact slab bookkeeping of the reference (Time/Slab.hpp, Time.cpp:114-117)."""
        _check(self._lib.dgrhs_set_slab(self._h, ctypes.c_double(slab_start),
                                        ctypes.c_double(slab_end), int(steps_per_slab)))

    def self_start_substeps_left(self):
        n = ctypes.c_int(0)
        _check(self._lib.dgrhs_self_start_substeps_left(self._h, ctypes.byref(n)))
        return n.value

    def set_exponential_filter(self, enable: bool, alpha: float = 36.0, half_power: int = 64):
        _check(self._lib.dgrhs_set_exponential_filter(self._h, int(enable),
                                                      ctypes.c_double(alpha), half_power))

    def set_p_mortars(self, table):
        """[n, 4] rows (element, direction, neighbour's N, neighbour direction | perm << 3);
        the faces carry P_MORTAR in the neighbour table."""
        t = np.ascontiguousarray(table, dtype=np.int32).reshape(-1, 4)
        _check(self._lib.dgrhs_set_p_mortars(self._h, len(t), _ptr(t)))

    def p_mortar_transfer_from(self, src, src_slots, dst_faces):
        """the faces `src` packed into its halo slots -> this context's p-mortar faces"""
        a = np.ascontiguousarray(src_slots, dtype=np.int32)
        b = np.ascontiguousarray(dst_faces, dtype=np.int32)
        _check(self._lib.dgrhs_p_mortar_transfer(src._h, self._h, len(a), _ptr(a), _ptr(b)))

    def apply_exponential_filter(self):
        _check(self._lib.dgrhs_apply_exponential_filter(self._h))

    def set_mortars(self, mortars):
        """[n, 6] rows (coarse element, direction, fine element, direction, size_a,
        size_b); the faces involved carry HANGING in the neighbour table."""
        m = np.ascontiguousarray(mortars, dtype=np.int32).reshape(-1, 6)
        _check(self._lib.dgrhs_set_mortars(self._h, len(m), _ptr(m)))

    def set_demand_outgoing_char_speeds(self, enable=True):
        _check(self._lib.dgrhs_set_demand_outgoing_char_speeds(self._h, int(enable)))

    def check_outgoing_char_speeds(self):
        """Raises DgrhsError if a DemandOutgoingCharSpeeds face saw an ingoing
        characteristic speed since the check was enabled."""
        n = ctypes.c_longlong(0)
        mn = ctypes.c_double(0.0)
        _check(self._lib.dgrhs_check_outgoing_char_speeds(self._h, ctypes.byref(n),
                                                          ctypes.byref(mn)))

    def set_split_volume(self, variant):
        """0 default kernels, 1 context + streaming kernels (N <= 10), 2 pair-staged
        kernel also for N >= 10 (instead of the component-slot kernel)."""
        _check(self._lib.dgrhs_set_split_volume(self._h, int(variant)))

    def set_fused_update(self, enable: bool):
        _check(self._lib.dgrhs_set_fused_update(self._h, int(enable)))

    def take_steps(self, n):
        _check(self._lib.dgrhs_take_steps(self._h, n))

    def begin_substep(self) -> float:
        t = ctypes.c_double()
        _check(self._lib.dgrhs_begin_substep(self._h, ctypes.byref(t)))
        return t.value

    def end_substep(self) -> bool:
        done = ctypes.c_int()
        _check(self._lib.dgrhs_end_substep(self._h, ctypes.byref(done)))
        return bool(done.value)

    @property
    def time(self):
        return float(self._lib.dgrhs_time(self._h))

    @property
    def rhs_evaluations(self):
        return int(self._lib.dgrhs_rhs_evaluations(self._h))

    def time_kernels(self, reps=5, update_terms=3):
        """Mean ms per launch of (face kernel, volume kernel, stepper update,
        volume kernel with the update fused in, exponential filter pass or 0)."""
        ms = np.zeros(5)
        _check(self._lib.dgrhs_time_kernels(self._h, reps, update_terms, _ptr(ms)))
        return ms

    def gh_constraint_norms(self):
        """L2 norms of the (gauge, three-index, four-index) constraints."""
        out = np.zeros(3)
        _check(self._lib.dgrhs_gh_constraint_norms(self._h, _ptr(out)))
        return out

    def synchronize(self):
        _check(self._lib.dgrhs_synchronize(self._h))

    @property
    def stream(self):
        return self._lib.dgrhs_stream(self._h)

    @property
    def state_device_ptr(self):
        return self._lib.dgrhs_state_device_ptr(self._h)

    @property
    def padded_points(self):
        return int(self._lib.dgrhs_padded_points(self._h))
