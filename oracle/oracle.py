"""numpy half of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``spectre_b200`` never does.

It restates the host-side pieces of the reference's DG evolution path --
spectral matrices, Adams-Bashforth / Runge-Kutta stepping with the forward
self-start, the periodic Brick domain, analytic initial data and the parity
norms -- and wraps the C restatement in ``dg_oracle.c`` (heavy loops).
Citations are to the reference checkout (v2024.09.29), file:line.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False, optimized: bool = False) -> str:
    """Compile dg_oracle.c -> oracle/_build/liboracle.so (gcc, OpenMP).
    optimized=False: the checker (-O2, no FMA contraction, so that the operation
    order of the restatement is what the compiler emits); optimized=True: the
    timing build of bench.py's CPU arm (-O3 -march=native) in
    oracle/_build/liboracle_fast_<cpu>.so -- never used for parity."""
    name = "liboracle.so"
    if optimized:
        # -march=native code must not travel to a different CPU: key it by the flags
        import hashlib
        try:
            flags_line = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
        except (OSError, StopIteration):
            flags_line = "unknown"
        name = f"liboracle_fast_{hashlib.sha1(flags_line.encode()).hexdigest()[:10]}.so"
    flags = ["-O3", "-march=native"] if optimized else ["-O2", "-ffp-contract=off"]
    out = os.path.join(_HERE, "_build", name)
    src = os.path.join(_HERE, "dg_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-std=c11", *flags, "-fopenmp", "-fPIC", "-shared", "-o", out, src, "-lm"]
        )
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def use_optimized_build():
    """Switch this process to the -O3 -march=native build (bench.py CPU arm only)."""
    global _LIB
    _LIB = ctypes.CDLL(build(optimized=True))
    _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------------------
# Spectral: NumericalAlgorithms/Spectral/Legendre.cpp:187-232 (LGL nodes and
# weights, Kopriva Alg. 25), Spectral.cpp:84-104 (barycentric weights, Kopriva
# Alg. 30), Spectral.cpp:431-445 (differentiation matrix).
# ---------------------------------------------------------------------------
def _q_and_L(poly_degree: int, x: float):
    # Legendre.cpp:160-184 (EvaluateQandL): three-term recurrence up to
    # L_{N+1}; q = L_{N+1} - L_{N-1}
    L_nm2, L_nm1 = 1.0, x
    L_n = x
    for k in range(2, poly_degree + 1):
        L_n = ((2.0 * k - 1.0) * x * L_nm1 - (k - 1.0) * L_nm2) / k
        L_nm2, L_nm1 = L_nm1, L_n
    k = poly_degree + 1
    L_np1 = ((2.0 * k - 1.0) * x * L_n - (k - 1.0) * L_nm2) / k
    return L_np1 - L_nm2, L_n


def lgl_points_and_weights(num_points: int):
    N = num_points - 1
    x = np.zeros(num_points)
    w = np.zeros(num_points)
    if N == 1:
        x[:] = [-1.0, 1.0]
        w[:] = 1.0
        return x, w
    x[0], x[N] = -1.0, 1.0
    w[0] = w[N] = 2.0 / (N * (N + 1.0))
    for j in range(1, (N + 1) // 2):
        lo = -math.cos((j - 0.25) * math.pi / N - 0.375 / (N * math.pi * (j - 0.25)))
        hi = -math.cos((j + 0.75) * math.pi / N - 0.375 / (N * math.pi * (j + 0.75)))
        # bracketing root find of q to full double precision (the reference
        # uses TOMS748 with abs tol 6e-16)
        flo = _q_and_L(N, lo)[0]
        fhi = _q_and_L(N, hi)[0]
        assert flo * fhi < 0.0
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if mid == lo or mid == hi:
                break
            fm = _q_and_L(N, mid)[0]
            if fm == 0.0:
                lo = hi = mid
                break
            if (fm < 0) == (flo < 0):
                lo, flo = mid, fm
            else:
                hi, fhi = mid, fm
        root = 0.5 * (lo + hi)
        # polish with a secant step (keeps |q| minimal)
        L = _q_and_L(N, root)[1]
        x[j] = root
        x[N - j] = -root
        w[j] = w[N - j] = 2.0 / (N * (N + 1.0) * L * L)
    if N % 2 == 0:
        L = _q_and_L(N, 0.0)[1]
        x[N // 2] = 0.0
        w[N // 2] = 2.0 / (N * (N + 1.0) * L * L)
    return x, w


def barycentric_weights(x):
    n = len(x)
    bw = np.ones(n)
    for j in range(1, n):
        for k in range(j):
            bw[k] *= x[k] - x[j]
            bw[j] *= x[j] - x[k]
    return 1.0 / bw


def differentiation_matrix(num_points: int):
    """D[i, j]; row-major (the reference stores it column-major)."""
    x, _ = lgl_points_and_weights(num_points)
    bw = barycentric_weights(x)
    D = np.zeros((num_points, num_points))
    for i in range(num_points):
        diag = 0.0
        for j in range(num_points):
            if i != j:
                D[i, j] = bw[j] / (bw[i] * (x[i] - x[j]))
                diag -= D[i, j]
        D[i, i] = diag
    return D


def exponential_filter_matrix(num_points: int, alpha: float, half_power: int):
    """Spectral::filtering::exponential_filter (Spectral/Filtering.cpp:20-32):
    V diag(exp(-alpha (i/(N-1))^(2 half_power))) V^-1 with the Legendre
    Vandermonde matrix at the LGL points (Spectral.cpp:498-523; the inverse is
    numerical there too)."""
    x, _ = lgl_points_and_weights(num_points)
    V = np.zeros((num_points, num_points))
    for j in range(num_points):
        c = np.zeros(j + 1)
        c[j] = 1.0
        V[:, j] = np.polynomial.legendre.legval(x, c)
    order = num_points - 1.0
    lam = np.exp(-alpha * (np.arange(num_points) / order) ** (2 * half_power))
    return V @ np.diag(lam) @ np.linalg.inv(V)


def apply_filter(N, u, F):
    """apply_matrices(u, {F, F, F}) on every component: u [..., n]."""
    shp = u.shape
    v = u.reshape(-1, N, N, N)  # [.., k, j, i]
    v = np.einsum("im,...kjm->...kji", F, v)
    v = np.einsum("jm,...kmi->...kji", F, v)
    v = np.einsum("km,...mji->...kji", F, v)
    return v.reshape(shp)


# ---------------------------------------------------------------------------
# Adams-Bashforth coefficients: Time/TimeSteppers/AdamsCoefficients.cpp:13-42
# (constant-step table), :75-117 (variable_coefficients), AdamsCoefficients.hpp
# :64-104 (selection logic).
# ---------------------------------------------------------------------------
_AB_CONST = {
    1: [1.0],
    2: [-0.5, 1.5],
    3: [5.0 / 12.0, -4.0 / 3.0, 23.0 / 12.0],
    4: [-3.0 / 8.0, 37.0 / 24.0, -59.0 / 24.0, 55.0 / 24.0],
    5: [251.0 / 720.0, -637.0 / 360.0, 109.0 / 30.0, -1387.0 / 360.0, 1901.0 / 720.0],
    6: [-95.0 / 288.0, 959.0 / 480.0, -3649.0 / 720.0, 4991.0 / 720.0, -2641.0 / 480.0,
        4277.0 / 1440.0],
    7: [19087.0 / 60480.0, -5603.0 / 2520.0, 135713.0 / 20160.0, -10754.0 / 945.0,
        235183.0 / 20160.0, -18637.0 / 2520.0, 198721.0 / 60480.0],
    8: [-5257.0 / 17280.0, 32863.0 / 13440.0, -115747.0 / 13440.0, 2102243.0 / 120960.0,
        -296053.0 / 13440.0, 242653.0 / 13440.0, -1152169.0 / 120960.0, 16083.0 / 4480.0],
}


def variable_coefficients(control_times, step_start, step_end):
    ct = [t - step_start for t in control_times]
    order = len(ct)
    result = []
    for j in range(order):
        poly = [0.0] * order
        poly[0] = 1.0
        for m in range(order):
            if m == j:
                continue
            denom = 1.0 / (ct[j] - ct[m])
            i = m + 1 if m < j else m
            while i > 0:
                poly[i] = (poly[i - 1] - poly[i] * ct[m]) * denom
                i -= 1
            poly[0] *= -ct[m] * denom
        for m in range(order):
            poly[m] /= m + 1
        dt = step_end - step_start
        val = 0.0
        for c in reversed(poly):  # evaluate_polynomial: Horner
            val = val * dt + c
        result.append(dt * val)
    return result


def ab_coefficients(times, step_start, step_end):
    """times: history times oldest first (floats or Fractions)."""
    if not times:
        return []
    step_size = float(step_end - step_start)
    constant = True
    control = [0.0]
    prev = times[0]
    for t in times[1:]:
        this_step = float(t - prev)
        control.append(control[-1] + this_step)
        if constant and abs(this_step - step_size) > 4.0 * np.finfo(float).eps * max(
            abs(float(t)), abs(step_size), 1e-300
        ):
            constant = False
        prev = t
    if constant and step_start == prev:
        return [c * step_size for c in _AB_CONST[len(control)]]
    return variable_coefficients(
        control, control[-1] + float(step_start - prev), control[-1] + float(step_end - prev)
    )


# ---------------------------------------------------------------------------
# Domain: periodic Brick (Domain/Creators/Rectilinear.cpp + Affine map).
# Element order: x fastest (index = ix + nx*(iy + ny*iz)).
# ---------------------------------------------------------------------------
class Brick:
    def __init__(self, lower, upper, refinement, N, periodic=True):
        self.lower = np.asarray(lower, float)
        self.upper = np.asarray(upper, float)
        self.ne = [2 ** r for r in refinement]
        self.N = N
        self.nelem = self.ne[0] * self.ne[1] * self.ne[2]
        self.n = N ** 3
        self.periodic = periodic
        x1, _ = lgl_points_and_weights(N)
        self.xi = x1

    def element_bounds(self, e):
        nx, ny, nz = self.ne
        idx = (e % nx, (e // nx) % ny, e // (nx * ny))
        h = (self.upper - self.lower) / np.asarray(self.ne)
        lo = self.lower + h * np.asarray(idx)
        return lo, lo + h

    def coords(self):
        """[nelem, 3, n] inertial coordinates."""
        N, n = self.N, self.n
        out = np.zeros((self.nelem, 3, n))
        i = np.arange(n) % N
        j = (np.arange(n) // N) % N
        k = np.arange(n) // (N * N)
        for e in range(self.nelem):
            lo, hi = self.element_bounds(e)
            for d, idx in enumerate((i, j, k)):
                # Affine: x = (hi-lo)/2 * xi + (hi+lo)/2  (CoordinateMaps/Affine.cpp)
                out[e, d] = 0.5 * (hi[d] - lo[d]) * self.xi[idx] + 0.5 * (hi[d] + lo[d])
        return out

    def inverse_jacobian(self):
        """[nelem, 9, n], comp = ihat + 3*i."""
        out = np.zeros((self.nelem, 9, self.n))
        for e in range(self.nelem):
            lo, hi = self.element_bounds(e)
            for d in range(3):
                out[e, d + 3 * d] = 2.0 / (hi[d] - lo[d])
        return out

    def neighbors(self):
        nx, ny, nz = self.ne
        nbr = np.full((self.nelem, 6), -1, dtype=np.int32)
        for e in range(self.nelem):
            ix, iy, iz = e % nx, (e // nx) % ny, e // (nx * ny)
            for d in range(6):
                dim, side = d // 2, d % 2
                c = [ix, iy, iz]
                c[dim] += 1 if side else -1
                ext = [nx, ny, nz][dim]
                if c[dim] < 0 or c[dim] >= ext:
                    if not self.periodic:
                        continue
                    c[dim] %= ext
                nbr[e, d] = c[0] + nx * (c[1] + ny * c[2])
        return nbr


# ---------------------------------------------------------------------------
# Analytic data
# ---------------------------------------------------------------------------
def plane_wave(x, t, k=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0), amp=1.0, wavenumber=1.0,
               phase=0.0):
    """PointwiseFunctions/AnalyticSolutions/WaveEquation/PlaneWave.cpp:56-119
    with a MathFunctions::Sinusoid profile.  x: [..., 3, n] -> u [..., 5, n]."""
    k = np.asarray(k, float)
    omega = math.sqrt(float(np.dot(k, k)))
    u_arg = sum(k[i] * (x[..., i, :] - center[i]) for i in range(3)) - omega * t
    prof = amp * np.sin(wavenumber * u_arg + phase)
    dprof = amp * wavenumber * np.cos(wavenumber * u_arg + phase)
    out = np.zeros(x.shape[:-2] + (5, x.shape[-1]))
    out[..., 0, :] = prof
    out[..., 1, :] = omega * dprof  # Pi = -dpsi/dt
    for i in range(3):
        out[..., 2 + i, :] = k[i] * dprof
    return out


def sym4(a, b):
    if a > b:
        a, b = b, a
    return a * 4 - a * (a - 1) // 2 + (b - a)


def gh_vars_from_metric(g, dtg, dg):
    """g: [4,4,...], dtg: [4,4,...], dg: [3,4,4,...] -> u [50, ...].
    Phi_iab = d_i g_ab, Pi_ab = -(d_t g_ab - shift^i Phi_iab)/lapse
    (GeneralizedHarmonic/{Phi.cpp:25-48,Pi.cpp:26-55} composed with
    SpacetimeMetric.cpp; identical up to rounding)."""
    gam = g[1:, 1:]
    inv_gam = np.linalg.inv(np.moveaxis(gam, (0, 1), (-2, -1)))
    inv_gam = np.moveaxis(inv_gam, (-2, -1), (0, 1))
    shift = np.einsum("ij...,j...->i...", inv_gam, g[1:, 0])
    lapse = np.sqrt(-g[0, 0] + np.einsum("i...,i...->...", shift, g[1:, 0]))
    pi = -(dtg - np.einsum("i...,iab...->ab...", shift, dg)) / lapse
    u = np.zeros((50,) + g.shape[2:])
    for a in range(4):
        for b in range(a, 4):
            s = sym4(a, b)
            u[s] = g[a, b]
            u[10 + s] = pi[a, b]
            for i in range(3):
                u[20 + i + 3 * s] = dg[i, a, b]
    return u


def gauge_wave_metric(x, t, amplitude=0.1, wavelength=1.0):
    """AnalyticSolutions/GeneralRelativity/GaugeWave.hpp:34-50:
    ds^2 = -H dt^2 + H dx^2 + dy^2 + dz^2, H = 1 - A sin(2 pi (x-t)/d).
    x: [3, n]."""
    omega = 2.0 * np.pi / wavelength
    H = 1.0 - amplitude * np.sin(omega * (x[0] - t))
    dH = -omega * amplitude * np.cos(omega * (x[0] - t))
    n = x.shape[-1]
    g = np.zeros((4, 4, n))
    dtg = np.zeros((4, 4, n))
    dg = np.zeros((3, 4, 4, n))
    g[0, 0] = -H
    g[1, 1] = H
    g[2, 2] = 1.0
    g[3, 3] = 1.0
    dtg[0, 0] = dH
    dtg[1, 1] = -dH
    dg[0, 0, 0] = -dH
    dg[0, 1, 1] = dH
    return g, dtg, dg


def kerr_schild_metric(x, mass=1.0, center=(0.0, 0.0, 0.0)):
    """Non-spinning Kerr-Schild (AnalyticSolutions/GeneralRelativity/
    KerrSchild.hpp:40-200 with a=0): g = eta + 2 H l l, H = M/r, l = (1, x/r)."""
    n = x.shape[-1]
    xc = np.stack([x[i] - center[i] for i in range(3)])
    r = np.sqrt(np.sum(xc * xc, axis=0))
    H = mass / r
    l = np.zeros((4, n))
    l[0] = 1.0
    l[1:] = xc / r
    dH = -mass * xc / r ** 3  # [3, n]
    dl = np.zeros((3, 4, n))
    for i in range(3):
        for j in range(3):
            dl[i, 1 + j] = ((1.0 if i == j else 0.0) - xc[i] * xc[j] / r ** 2) / r
    eta = np.diag([-1.0, 1.0, 1.0, 1.0])
    g = eta[:, :, None] + 2.0 * H * l[:, None] * l[None, :]
    dg = np.zeros((3, 4, 4, n))
    for i in range(3):
        dg[i] = 2.0 * dH[i] * l[:, None] * l[None, :] + 2.0 * H * (
            dl[i][:, None] * l[None, :] + l[:, None] * dl[i][None, :]
        )
    dtg = np.zeros((4, 4, n))
    return g, dtg, dg


def gaussian_plus_constant(x, constant, amplitude, width, center):
    """ConstraintDamping/GaussianPlusConstant.cpp: C + A exp(-|x-x0|^2/w^2)."""
    r2 = sum((x[..., i, :] - center[i]) ** 2 for i in range(3))
    return constant + amplitude * np.exp(-r2 / width ** 2)


# ---------------------------------------------------------------------------
# C wrappers
# ---------------------------------------------------------------------------
GAUGE_HARMONIC = np.array([0.0, 0, 0, 0, 0, 0, 0, 0])
GAUGE_GIVEN = np.array([1.0, 0, 0, 0, 0, 0, 0, 0])


def partial_derivatives(N, u, invjac):
    """u [C, n], invjac [9, n] -> du [3C, n]"""
    C = u.shape[0]
    D = _c(differentiation_matrix(N))
    u, invjac = _c(u), _c(invjac)
    du = np.zeros((3 * C, N ** 3))
    lib().orc_partial_derivatives(N, C, _p(D), _p(u), _p(invjac), _p(du))
    return du


def sw_time_derivative(u, du, gamma2):
    n = u.shape[1]
    dt = np.zeros((5, n))
    u, du, gamma2 = _c(u), _c(du), _c(gamma2)
    lib().orc_sw_time_derivative(n, _p(u), _p(du), _p(gamma2), _p(dt))
    return dt


def gh_time_derivative(u, du, gamma0, gamma1, gamma2, gauge_params=GAUGE_HARMONIC, H=None,
                       dH=None, coords=None):
    n = u.shape[1]
    dt = np.zeros((50, n))
    args = [_c(a) for a in (u, du, gamma0, gamma1, gamma2, gauge_params)]
    H = _c(H) if H is not None else np.zeros((4, n))
    dH = _c(dH) if dH is not None else np.zeros((16, n))
    coords = _c(coords) if coords is not None else None
    lib().orc_gh_time_derivative(n, *[_p(a) for a in args], _p(H), _p(dH), _p(coords), _p(dt))
    return dt


def gh_geometry(u):
    n = u.shape[1]
    out = np.zeros((21, n))
    u = _c(u)
    lib().orc_gh_geometry(n, _p(u), _p(out))
    return {"lapse": out[0], "shift": out[1:4], "inv_gamma": out[4:10], "inv_g": out[10:20],
            "det_gamma": out[20]}


def gh_package_data_moving(u, gamma1, gamma2, lapse, shift, n_lo, n_up, ndotv):
    """GH dg_package_data on f face points with normal_dot_mesh_velocity
    (UpwindPenalty.cpp:36-158); u [50, f], shift / normals [3, f]; packaged [134, f]."""
    f = u.shape[1]
    out = np.zeros((134, f))
    a = [_c(x) for x in (u, gamma1, gamma2, lapse, shift, n_lo, n_up, ndotv)]
    lib().orc_gh_package_data_moving(f, *[_p(x) for x in a], _p(out))
    return out


def sw_package_data_moving(u, gamma2, normal, ndotv):
    """ScalarWave dg_package_data with normal_dot_mesh_velocity (UpwindPenalty.cpp:36-108);
    u [5, f], normal [3, f]; packaged [16, f]."""
    f = u.shape[1]
    out = np.zeros((16, f))
    a = [_c(x) for x in (u, gamma2, normal, ndotv)]
    lib().orc_sw_package_data_moving(f, *[_p(x) for x in a], _p(out))
    return out


def dg_rhs(system, N, u, invjac, static_fields, nbr, gauge_params=GAUGE_HARMONIC, coords=None,
           volume_only=False, ext_u=None, nbr_dir=None, face_perm=None, mortars=None,
           mesh_velocity=None):
    """u [nelem, C, n]; returns dt_u of the same shape.  ext_u [nslots, C, f]:
    exterior states of ghost boundary conditions (nbr <= -2 -> slot -(nbr+2)).
    nbr_dir / face_perm [nelem, 6]: orientation of non-aligned neighbours (the
    neighbour's direction touching the face and the face-point permutation code,
    see orc_dg_rhs_oriented).  mortars [n, 6]: non-conforming (2:1) mortars
    (coarse element, direction, fine element, direction, size_a, size_b); the
    faces involved carry HANGING in nbr (see orc_dg_rhs_mortars).
    mesh_velocity [nelem, 3, n]: inertial mesh velocity of a moving mesh (VolumeTermsImpl.tpp:
    155-235, GH TimeDerivative.cpp:237-300,372-378, normal_dot_mesh_velocity in dg_package_data;
    conforming faces and ghost boundary conditions only)."""
    nelem = u.shape[0]
    if mesh_velocity is not None:
        assert not volume_only and (mortars is None or len(mortars) == 0)
        mv = _c(mesh_velocity)
        lib().orc_set_mesh_velocity(_p(mv))
        try:
            return dg_rhs(system, N, u, invjac, static_fields, nbr, gauge_params, coords,
                          volume_only, ext_u, nbr_dir, face_perm, mortars)
        finally:
            lib().orc_set_mesh_velocity(None)
    D = _c(differentiation_matrix(N))
    u, invjac, static_fields = _c(u), _c(invjac), _c(static_fields)
    gp = _c(gauge_params)
    nbr = np.ascontiguousarray(nbr, dtype=np.int32)
    coords = _c(coords) if coords is not None else None
    dt = np.zeros_like(u)
    if volume_only:
        lib().orc_dg_volume(system, N, nelem, _p(D), _p(u), _p(invjac), _p(static_fields),
                            _p(coords), _p(gp), _p(dt))
    else:
        ext = _c(ext_u) if ext_u is not None else None
        nf = None
        if nbr_dir is not None:
            nf = np.ascontiguousarray(np.asarray(nbr_dir) | (np.asarray(face_perm) << 3),
                                      dtype=np.int32)
        nm, mt, P, R = 0, None, None, None
        if mortars is not None and len(mortars):
            mt = np.ascontiguousarray(mortars, dtype=np.int32)
            nm = len(mt)
            P = _c(np.stack([np.eye(N)] + [projection_matrix_parent_to_child(N, N, sz)
                                           for sz in (MORTAR_LOWER_HALF, MORTAR_UPPER_HALF)]))
            R = _c(np.stack([np.eye(N)] + [projection_matrix_child_to_parent(N, N, sz)
                                           for sz in (MORTAR_LOWER_HALF, MORTAR_UPPER_HALF)]))
        lib().orc_dg_rhs_mortars(system, N, nelem, _p(D), _p(u), _p(invjac),
                                 _p(static_fields), _p(coords), _p(nbr), _p(nf), _p(gp),
                                 _p(ext), nm, _p(mt), _p(P), _p(R), _p(dt))
        if (nbr == BJORHUS).any() or (nbr == BJORHUS_PHYSICAL).any():
            dt += bjorhus_corrections(N, u, invjac, static_fields, coords, nbr, gauge_params)
    return dt


def bjorhus_corrections(N, u, invjac, static_fields, coords, nbr, gauge_params=GAUGE_HARMONIC):
    """TimeDerivative-type boundary condition ConstraintPreservingBjorhus on the
    faces marked BJORHUS (BoundaryConditionsImpl.hpp:566-670): the volume time
    derivative and the volume partial derivatives are sliced to the face, the
    condition returns corrections to dt(g, Pi, Phi) which are added on the face
    points (no lifting).  Returns the array to add to the right-hand side."""
    from . import bjorhus as bj
    nelem, n = u.shape[0], N ** 3
    given = static_fields.shape[1] >= 23
    vol = dg_rhs(1, N, u, invjac, static_fields, nbr, gauge_params, coords, volume_only=True)
    out = np.zeros_like(u)
    unpack = lambda v: np.array([[v[sym4(a, b)] for b in range(4)] for a in range(4)])
    for e in range(nelem):
        faces = [d for d in range(6) if nbr[e, d] in (BJORHUS, BJORHUS_PHYSICAL)]
        if not faces:
            continue
        du = np.asarray(partial_derivatives(N, u[e], invjac[e])).reshape(50, 3, n)
        geo = gh_geometry(u[e])
        for d in faces:
            dim, sign = d // 2, (1.0 if d % 2 else -1.0)
            a, b = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
            fixed = N - 1 if d % 2 else 0
            pts = [fixed + N * (a + N * b), a + N * (fixed + N * b), a + N * (b + N * fixed)][dim]
            for p in pts.ravel():
                ig = np.zeros((3, 3))
                c = 0
                for i in range(3):
                    for j in range(i, 3):
                        ig[i, j] = ig[j, i] = geo["inv_gamma"][c][p]
                        c += 1
                unnorm = sign * np.array([invjac[e][dim + 3 * i][p] for i in range(3)])
                n_lo = unnorm / np.sqrt(unnorm @ ig @ unnorm)
                lapse, shift = geo["lapse"][p], geo["shift"][:, p]
                ipsi = unpack(geo["inv_g"][:, p])
                t_up = np.concatenate([[1.0 / lapse], -shift / lapse])
                g = unpack(u[e][0:10, p])
                pi = unpack(u[e][10:20, p])
                tens3 = lambda arr: np.array([unpack(np.array([arr[20 + m + 3 * s]
                                                               for s in range(10)]))
                                              for m in range(3)])
                phi = tens3(u[e][:, p])
                dt_g, dt_pi = unpack(vol[e][0:10, p]), unpack(vol[e][10:20, p])
                dt_phi = tens3(vol[e][:, p])
                d_g = np.array([unpack(du[0:10, i, p]) for i in range(3)])
                d_pi = np.array([unpack(du[10:20, i, p]) for i in range(3)])
                d_phi = np.array([[unpack(np.array([du[20 + m + 3 * s, i, p] for s in range(10)]))
                                   for m in range(3)] for i in range(3)])
                if gauge_params[0] == 2.0:   # DampedHarmonic: evaluate it at the point
                    H, dH = np.zeros(4), np.zeros((4, 4))
                    P_ = lambda a_: np.ascontiguousarray(a_, dtype=np.float64)
                    gg, pp, ff = P_(g), P_(pi), P_(phi)
                    xx, gp_ = P_(coords[e][:, p]), P_(gauge_params)
                    lib().orc_damped_harmonic(_p(gg), _p(pp), _p(ff), _p(xx), _p(gp_), _p(H),
                                              _p(dH))
                elif given:
                    H = static_fields[e][3:7, p]
                    dH = np.array([[static_fields[e][7 + aa + 4 * bb, p] for bb in range(4)]
                                   for aa in range(4)])
                else:
                    H, dH = np.zeros(4), np.zeros((4, 4))
                cg, cp, cph = bj.bjorhus_constraint_preserving(
                    n_lo, g, pi, phi, coords[e][:, p], static_fields[e][1, p],
                    static_fields[e][2, p], lapse, shift, ipsi, t_up, d_g - phi, H, dH,
                    dt_g, dt_pi, dt_phi, d_pi, d_phi, physical=nbr[e, d] == BJORHUS_PHYSICAL)
                for s_, (aa, bb) in enumerate([(x, y) for x in range(4) for y in range(x, 4)]):
                    out[e][s_, p] += cg[aa, bb]
                    out[e][10 + s_, p] += cp[aa, bb]
                    for m in range(3):
                        out[e][20 + m + 3 * s_, p] += cph[m, aa, bb]
    return out


def gh_characteristic_speeds(gamma1, lapse, shift, unit_normal_one_form):
    """gh::characteristic_speeds without mesh velocity (GeneralizedHarmonic/
    Characteristics.cpp:24-40): lambda(VSpacetimeMetric, VZero, VPlus, VMinus)."""
    sdn = np.dot(shift, unit_normal_one_form)
    return np.array([-(1.0 + gamma1) * sdn, -sdn, -sdn + lapse, -sdn - lapse])


def demand_outgoing_char_speeds(N, u, invjac, gamma1, nbr):
    """DemandOutgoingCharSpeeds::dg_demand_outgoing_char_speeds (GeneralizedHarmonic/
    BoundaryConditions/DemandOutgoingCharSpeeds.cpp:37-76) on every face with
    nbr == -1: returns (number of face points with an ingoing speed, most negative
    speed).  Normal: NormalCovectorAndMagnitude.hpp:47-92."""
    n_bad, worst = 0, 0.0
    for e in range(u.shape[0]):
        for d in range(6):
            if nbr[e, d] != -1:
                continue
            dim, sign = d // 2, (1.0 if d % 2 else -1.0)
            a, b = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
            fixed = N - 1 if d % 2 else 0
            pts = ([fixed + N * (a + N * b), a + N * (fixed + N * b), a + N * (b + N * fixed)][dim]
                   ).ravel()
            geo = gh_geometry(u[e][:, pts])
            for k, p in enumerate(pts):
                unnorm = sign * np.array([invjac[e][dim + 3 * i][p] for i in range(3)])
                ig = np.zeros((3, 3))
                c = 0
                for i in range(3):
                    for j in range(i, 3):
                        ig[i, j] = ig[j, i] = geo["inv_gamma"][c][k]
                        c += 1
                mag = np.sqrt(unnorm @ ig @ unnorm)
                lam = gh_characteristic_speeds(gamma1[e][p], geo["lapse"][k],
                                               geo["shift"][:, k], unnorm / mag)
                if lam.min() < 0.0:
                    n_bad += 1
                    worst = min(worst, lam.min())
    return n_bad, worst


def analytic_christoffel_gauge(N, u_analytic, invjac):
    """H_a = -Gamma_a of the analytic solution, d_i H_a by numerical
    differentiation, d_t H_a = 0 (GaugeSourceFunctions/AnalyticChristoffel.cpp
    :64-149).  u_analytic [50, n] -> H [4, n], dH [16, n] (d_a H_b at a+4b)."""
    n = u_analytic.shape[1]
    g = np.zeros((4, 4, n))
    pi = np.zeros((4, 4, n))
    phi = np.zeros((3, 4, 4, n))
    for a in range(4):
        for b in range(4):
            s = sym4(a, b)
            g[a, b] = u_analytic[s]
            pi[a, b] = u_analytic[10 + s]
            for i in range(3):
                phi[i, a, b] = u_analytic[20 + i + 3 * s]
    geo = gh_geometry(u_analytic)
    inv_g = np.zeros((4, 4, n))
    k = 0
    for a in range(4):
        for b in range(a, 4):
            inv_g[a, b] = inv_g[b, a] = geo["inv_g"][k]
            k += 1
    dag = np.zeros((4, 4, 4, n))
    dag[0] = -geo["lapse"] * pi + np.einsum("i...,iab...->ab...", geo["shift"], phi)
    dag[1:] = phi
    chr1 = 0.5 * (np.einsum("ijk...->kij...", dag) + np.einsum("jik...->kij...", dag) - dag)
    H = -np.einsum("abc...,bc...->a...", chr1, inv_g)
    dHs = partial_derivatives(N, H, invjac)  # [3*4, n], index 3*a + i
    dH = np.zeros((16, n))
    for b in range(4):
        for i in range(3):
            dH[(i + 1) + 4 * b] = dHs[3 * b + i]
    return H, dH


# ---------------------------------------------------------------------------
# Time stepping (GTS): AdamsBashforth.cpp:120-201, Rk3HesthavenSsp.cpp:55-81,
# self start Time/Actions/SelfStartActions.hpp:181-243,316-394 and the action
# order of Evolution/Executables/.../step_actions.
# Times are exact Fractions of the step (the reference uses rational Time).
# ---------------------------------------------------------------------------
# Butcher tableaus (substep times c, substep coefficients A, result b) of the
# reference's RungeKutta steppers: Rk3Owren.cpp:17-34, Rk3Kennedy.cpp:18-43,
# ClassicalRungeKutta4.cpp:24-49, DormandPrince5.cpp:21-50.
RK_TABLEAUS = {
    "Rk3Owren": ([12.0 / 23.0, 4.0 / 5.0],
                 [[12.0 / 23.0], [-68.0 / 375.0, 368.0 / 375.0]],
                 [31.0 / 144.0, 529.0 / 1152.0, 125.0 / 384.0]),
    "Rk3Kennedy": ([1767732205903.0 / 2027836641118.0, 3.0 / 5.0, 1.0],
                   [[1767732205903.0 / 2027836641118.0],
                    [5535828885825.0 / 10492691773637.0, 788022342437.0 / 10882634858940.0],
                    [6485989280629.0 / 16251701735622.0, -4246266847089.0 / 9704473918619.0,
                     10755448449292.0 / 10357097424841.0]],
                   [1471266399579.0 / 7840856788654.0, -4482444167858.0 / 7529755066697.0,
                    11266239266428.0 / 11593286722821.0, 1767732205903.0 / 4055673282236.0]),
    "RK4": ([1.0 / 2.0, 1.0 / 2.0, 1.0, 3.0 / 4.0],
            [[1.0 / 2.0], [0.0, 1.0 / 2.0], [0.0, 0.0, 1.0],
             [5.0 / 32.0, 7.0 / 32.0, 13.0 / 32.0, -1.0 / 32.0]],
            [1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0]),
    "DP5": ([1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0, 1.0],
            [[1.0 / 5.0], [3.0 / 40.0, 9.0 / 40.0], [44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0],
             [19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0],
             [9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0,
              -5103.0 / 18656.0],
             [35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0]],
            [35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0]),
}


class Evolution:
    def __init__(self, rhs, u0, t0, dt, stepper="AB3", post_update=None, slab=None):
        """rhs(u, t) -> dt_u.  stepper: 'AB<k>' or 'RK3' (Rk3HesthavenSsp).
        post_update(u) -> u: action after UpdateU in step_actions (the filter).
        slab = (start, end, steps_per_slab): times and step size formed as the reference's
        Slab / Time / TimeDelta / TimeStepId form them (Slab.hpp advance, Time.cpp:114-117,
        :127-129, TimeStepId.cpp:72-82) instead of t0 + k dt; t0 and dt are then ignored."""
        self.post = post_update if post_update is not None else (lambda v: v)
        self.rhs = rhs
        self.u = u0.copy()
        self.slab = slab
        if slab is not None:
            t0 = slab[0]
            dt = (slab[1] - slab[0]) * (1.0 / slab[2])
        self.t0 = t0
        self.dt = dt
        self.stepper = stepper
        self.step_index = 0  # completed full steps
        self.history = []  # list of (time_fraction, u or None, dt_u), oldest first
        self.rhs_evals = 0
        if stepper.startswith("AB"):
            self.order = int(stepper[2:])
            self._self_start()

    def _time(self, frac):
        if self.slab is None:
            return self.t0 + float(frac) * self.dt
        a, b, per_slab = self.slab
        f = Fraction(frac) / per_slab          # in slabs
        k = f.numerator // f.denominator
        f -= k
        for _ in range(k):                      # Slab::advance
            a, b = b, b + (b - a)
        g = 1 - f
        return (g.numerator / g.denominator) * a + (f.numerator / f.denominator) * b

    def _substep_time(self, n, c):
        """time of a substep at the fraction c (a double) of the step that starts at step n"""
        if c == 0.0:
            return self._time(n)
        if self.slab is None:
            return self.t0 + (float(n) + c) * self.dt
        return (1.0 - c) * self._time(n) + c * self._time(n + 1)

    def _eval(self, frac):
        self.rhs_evals += 1
        return self.rhs(self.u, self._time(frac))

    def _ab_update(self, order, step_start, step_end):
        hist = self.history[-order:]
        coefs = ab_coefficients_frac([h[0] for h in hist], step_start, step_end, self.dt)
        u = hist[-1][1].copy()
        for c, h in zip(coefs, hist):
            u += c * h[2]
        return u

    def _clean(self, order):
        while len(self.history) >= order:
            self.history.pop(0)
        if len(self.history) > 1:
            t, _, d = self.history[-2]
            self.history[-2] = (t, None, d)

    def _self_start(self):
        k = self.order
        if k == 1:
            return
        u_init = self.u.copy()
        h = Fraction(1, k)  # self-start step = dt / (values_needed + 1)
        for order in range(1, k):
            # reset to t0 (CheckForCompletion restores the initial value)
            self.u = u_init.copy()
            for s in range(order + 1):
                t = s * h
                d = self._eval(t)
                self.history.append((t, self.u.copy(), d))
                if s == order:
                    # step_unused: UpdateU skipped, history order was bumped
                    self._clean(order + 1)
                    break
                self.u = self.post(self._ab_update(order, t, t + h))
                self._clean(order)
        self.u = u_init.copy()

    def step(self):
        n = Fraction(self.step_index)
        if self.stepper.startswith("AB"):
            k = self.order
            d = self._eval(n)
            self.history.append((n, self.u.copy(), d))
            self.u = self.post(self._ab_update(k, n, n + 1))
            self._clean(k)
        elif self.stepper == "RK3":
            dt = self.dt
            u0 = self.u.copy()
            f0 = self._eval(n)
            self.u = self.post(u0 + dt * f0)
            self.rhs_evals += 1
            f1 = self.rhs(self.u, self._substep_time(n, 1.0))
            u1 = self.u.copy()
            self.u = self.post(0.25 * (3.0 * u0 + u1 + dt * f1))
            self.rhs_evals += 1
            f2 = self.rhs(self.u, self._substep_time(n, 0.5))
            u2 = self.u.copy()
            self.u = self.post((1.0 / 3.0) * (u0 + 2.0 * u2 + 2.0 * dt * f2))
        elif self.stepper in RK_TABLEAUS:
            # RungeKutta.cpp:69-122: u = u_start + dt sum_i coef_i f_i, the last
            # substep with the result coefficients; substep k > 0 at t + c[k-1] dt
            c, A, b = RK_TABLEAUS[self.stepper]
            nsub = len(b)
            u_start = self.u.copy()
            fs = []
            for k in range(nsub):
                self.rhs_evals += 1
                fs.append(self.rhs(self.u, self._substep_time(n, 0.0 if k == 0 else c[k - 1])))
                row = b if k == nsub - 1 else A[k]
                u = u_start.copy()
                for coef, f in zip(row, fs):
                    if coef != 0.0:
                        u += coef * self.dt * f
                self.u = self.post(u)
        else:
            raise ValueError(self.stepper)
        self.step_index += 1

    @property
    def time(self):
        return self._time(Fraction(self.step_index))


def ab_coefficients_frac(times, step_start, step_end, dt):
    """Same selection logic as AdamsCoefficients.hpp:64-104 with the history
    times given as exact Fractions of the step dt (the reference's Time is an
    exact rational of the slab)."""
    order = len(times)
    step = step_end - step_start
    uniform = all(b - a == step for a, b in zip(times[:-1], times[1:])) and \
        times[-1] == step_start
    if uniform:
        return [c * (float(step) * dt) for c in _AB_CONST[order]]
    control = [0.0]
    for a, b in zip(times[:-1], times[1:]):
        control.append(control[-1] + float(b - a) * dt)
    return variable_coefficients(control, control[-1] + float(step_start - times[-1]) * dt,
                                 control[-1] + float(step_end - times[-1]) * dt)


# ---------------------------------------------------------------------------
# Norms: ParallelAlgorithms/Events/ObserveNorms.hpp:60-80
# L2Norm with Components: Sum = sqrt( sum_points sum_comps v^2 / N_points )
# ---------------------------------------------------------------------------
def l2_norm(v):
    """v [nelem, ncomp, n] (independent components only, like the reference)."""
    npts = v.shape[0] * v.shape[2]
    return math.sqrt(float(np.sum(v * v)) / npts)


def four_index_constraint(d_phi):
    """C_i.. = eps_ijk d_j Phi_k.. (Constraints.cpp:1070-1100, python twin
    TestFunctions.py:295-300).  d_phi [3 (j), 3 (k), ...] -> [3, ...]."""
    eps = np.zeros((3, 3, 3))
    eps[0, 1, 2] = eps[1, 2, 0] = eps[2, 0, 1] = 1.0
    eps[0, 2, 1] = eps[2, 1, 0] = eps[1, 0, 2] = -1.0
    return np.einsum("ijk,jk...->i...", eps, d_phi)


def gh_constraint_norms(N, u, invjac, H=None):
    """L2 norms (ObserveNorms.hpp:60-80, Components: Sum) of the GH gauge
    constraint C_a = H_a + Gamma_a (Constraints.cpp:965-1000), the three-index
    constraint d_i g_ab - Phi_iab (:935-962) and the four-index constraint
    eps_ijk d_j Phi_kab (:1070-1100; python twin TestFunctions.py:295-300).
    u [nelem, 50, n], invjac [nelem, 9, n]."""
    nelem, _, n = u.shape
    sums = np.zeros(3)
    for e in range(nelem):
        Hg, _ = analytic_christoffel_gauge(N, u[e], invjac[e])  # = -Gamma_a
        gam = -Hg
        c1 = gam + (H[e] if H is not None else 0.0)
        sums[0] += np.sum(c1 * c1)
        du = partial_derivatives(N, u[e], invjac[e])
        for s in range(10):
            for i in range(3):
                c3 = du[3 * s + i] - u[e, 20 + i + 3 * s]
                sums[1] += np.sum(c3 * c3)
            dphi = np.zeros((3, 3, n))  # [j, k] = d_j Phi_k
            for j in range(3):
                for k in range(3):
                    dphi[j, k] = du[3 * (20 + k + 3 * s) + j]
            c4 = four_index_constraint(dphi)
            sums[2] += np.sum(c4 * c4)
    return np.sqrt(sums / (nelem * n))


# ---------------------------------------------------------------------------
# Non-conforming mortars: NumericalAlgorithms/Spectral/Projection.cpp and
# NumericalAlgorithms/DiscontinuousGalerkin/MortarHelpers.hpp:74-129.
# ChildSize / MortarSize codes: 0 Full, 1 LowerHalf, 2 UpperHalf.
# ---------------------------------------------------------------------------
MORTAR_FULL, MORTAR_LOWER_HALF, MORTAR_UPPER_HALF = 0, 1, 2
HANGING = -2 ** 31   # neighbour-table entry of a face that is handled by the mortar table
BJORHUS = -2 ** 31 + 1   # external face with ConstraintPreservingBjorhus (ConstraintPreserving)
BJORHUS_PHYSICAL = -2 ** 31 + 2   # ... Type ConstraintPreservingPhysical


def legendre_vandermonde(num_points):
    """modal_to_nodal_matrix (Spectral.cpp:499-516): V_ij = P_j(x_i)."""
    x, _ = lgl_points_and_weights(num_points)
    V = np.zeros((num_points, num_points))
    for j in range(num_points):
        V[:, j] = np.polynomial.legendre.legval(x, [0.0] * j + [1.0])
    return V


def interpolation_matrix(num_points, targets):
    """Spectral::interpolation_matrix (Spectral.cpp:625-680, Kopriva Alg. 32):
    barycentric Lagrange interpolation from the LGL points to `targets`."""
    x, _ = lgl_points_and_weights(num_points)
    bw = barycentric_weights(x)
    targets = np.atleast_1d(np.asarray(targets, float))
    M = np.zeros((len(targets), num_points))
    for k, t in enumerate(targets):
        match = [j for j in range(num_points)
                 if abs(t - x[j]) <= 1e-15 * max(1.0, abs(t), abs(x[j])) * 16]
        if match:
            M[k, match[0]] = 1.0
            continue
        row = bw / (t - x)
        M[k] = row / row.sum()
    return M


def projection_matrix_parent_to_child(n_parent, n_child, size):
    """Projection.cpp:279-362: interpolation from the parent's points to the
    child's points mapped into the parent interval (x, (x+1)/2 or (x-1)/2)."""
    xc, _ = lgl_points_and_weights(n_child)
    t = {MORTAR_FULL: xc, MORTAR_UPPER_HALF: 0.5 * (xc + 1.0),
         MORTAR_LOWER_HALF: 0.5 * (xc - 1.0)}[size]
    return interpolation_matrix(n_parent, t)


def _spectral_transformation(large_index, small_index):
    """Projection.cpp:158-186: the (large, small) entry of the map from the
    Legendre modes on the half interval to the modes on the whole interval."""
    assert large_index >= small_index
    result = 1.0
    for i in range((large_index - small_index) // 2, 0, -1):
        result = 1.0 - result * float(
            (large_index + small_index + 3 - 2 * i) * (large_index + small_index + 2 - 2 * i) *
            (large_index - small_index + 2 - 2 * i) * (large_index - small_index + 1 - 2 * i)) / \
            float(2 * i * (2 * large_index + 1 - 2 * i) * (large_index + 2 - 2 * i) *
                  (large_index + 1 - 2 * i))
    for i in range(1, large_index - small_index + 1):
        result *= 1.0 + float(large_index + small_index + 1) / i
    result /= 2.0 ** (large_index + 1)
    return result


def projection_matrix_child_to_parent(n_child, n_parent, size):
    """Projection.cpp:57-262 (operand not massive): the L2 projection from the
    child (mortar) to the parent (element face), done in modal space."""
    assert n_parent <= n_child
    V_el = legendre_vandermonde(n_parent)
    Vinv_mortar = np.linalg.inv(legendre_vandermonde(n_child))
    if size == MORTAR_FULL:
        # truncation of the modes
        return V_el @ Vinv_mortar[:n_parent, :]
    if size == MORTAR_UPPER_HALF:
        temp = np.zeros((n_parent, n_parent))
        for j in range(n_parent):
            for k in range(j, n_parent):
                temp[:, j] += V_el[:, k] * _spectral_transformation(k, j)
        return temp @ Vinv_mortar[:n_parent, :]
    upper = projection_matrix_child_to_parent(n_child, n_parent, MORTAR_UPPER_HALF)
    return upper[::-1, ::-1].copy()


# ---------------------------------------------------------------------------
# p-refinement: mortars between elements with different numbers of grid points
# (dg::mortar_mesh = the larger extents, MortarHelpers.cpp:22-49; project_to_mortar /
# project_from_mortar, MortarHelpers.hpp:74-129; ApplyBoundaryCorrections.hpp:286-380)
# ---------------------------------------------------------------------------
P_MORTAR = -2 ** 31 + 3


def face_point_indices(N, d):
    """volume indices of the points of face d, face index q = a + N b (a the first
    remaining dimension)"""
    dim, fixed = d // 2, (N - 1 if d % 2 else 0)
    a, b = np.meshgrid(np.arange(N), np.arange(N), indexing="xy")
    a, b = a.ravel(), b.ravel()
    if dim == 0:
        return fixed + N * (a + N * b)
    if dim == 1:
        return a + N * (fixed + N * b)
    return a + N * (b + N * fixed)


def face_packaged_data(system, N, u_e, invjac_e, static_e, d):
    """InternalMortarDataImpl.hpp:180-320 for one face: slice, unit normal
    (NormalCovectorAndMagnitude.hpp:47-92), dg_package_data.  Returns (packaged [PK, f],
    magnitude of the unnormalised normal [f])."""
    p = face_point_indices(N, d)
    f = N * N
    sign = 1.0 if d % 2 else -1.0
    unn = np.stack([sign * invjac_e[d // 2 + 3 * i][p] for i in range(3)])
    uf = _c(u_e[:, p])
    L = lib()
    if system == 0:
        mag = np.sqrt(unn[0] * unn[0] + unn[1] * unn[1] + unn[2] * unn[2])
        n_lo = _c(unn / mag)
        out = np.zeros((16, f))
        L.orc_sw_package_data(f, _p(uf), _p(_c(static_e[0][p])), _p(n_lo), _p(out))
        return out, mag
    geo = gh_geometry(uf)
    ig = geo["inv_gamma"]          # 00 01 02 11 12 22
    idx = [[0, 1, 2], [1, 3, 4], [2, 4, 5]]
    n_up = np.stack([ig[idx[i][0]] * unn[0] + ig[idx[i][1]] * unn[1] + ig[idx[i][2]] * unn[2]
                     for i in range(3)])
    mag = np.sqrt(n_up[0] * unn[0] + n_up[1] * unn[1] + n_up[2] * unn[2])
    n_lo, n_up = _c(unn / mag), _c(n_up / mag)
    out = np.zeros((134, f))
    L.orc_gh_package_data(f, _p(uf), _p(_c(static_e[1][p])), _p(_c(static_e[2][p])),
                          _p(_c(geo["lapse"])), _p(_c(geo["shift"])), _p(n_lo), _p(n_up), _p(out))
    return out, mag


def orient_face_map(n, perm):
    """index map q -> q' of orient_variables_on_slice on an n x n face: bit 0 swaps the face
    coordinates, bits 1 / 2 flip the first / second coordinate of the target frame"""
    a, b = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    a, b = a.ravel(), b.ravel()
    ta, tb = (b, a) if perm & 1 else (a, b)
    if perm & 2:
        ta = n - 1 - ta
    if perm & 4:
        tb = n - 1 - tb
    return ta + n * tb


def _apply_face_matrices(x, Ma, Mb):
    """apply_matrices on [C, nb, na] data: first face dimension, then the second"""
    if Ma is not None:
        x = np.einsum("ta,cba->cbt", Ma, x)
    if Mb is not None:
        x = np.einsum("tb,cba->cta", Mb, x)
    return x


def dg_rhs_p_refined(system, classes, links, gauge_params=GAUGE_HARMONIC):
    """Right-hand side of a domain whose elements come in classes with different N.
    classes: list of dicts N, u [E, C, n], invjac, static, nbr (faces to another class marked
    P_MORTAR; optional ext_u, nbr_dir, face_perm as in dg_rhs for the faces inside a class).  links: (class_a, element_a, direction_a, class_b, element_b, direction_b,
    perm) -- perm takes a face point of a to the same point in b's frame.
    Returns the list of dt_u per class."""
    L = lib()
    out = []
    for cl in classes:
        nbr = np.where(cl["nbr"] == P_MORTAR, -1, cl["nbr"]).astype(np.int32)
        out.append(dg_rhs(system, cl["N"], cl["u"], cl["invjac"], cl["static"], nbr,
                          gauge_params=gauge_params, ext_u=cl.get("ext_u"),
                          nbr_dir=cl.get("nbr_dir"), face_perm=cl.get("face_perm")))
    C, PK = (5, 16) if system == 0 else (50, 134)
    for (ca, ea, da, cb, eb, db, perm) in links:
        A, B = classes[ca], classes[cb]
        NA, NB = A["N"], B["N"]
        NM = max(NA, NB)
        side = []
        for cl, e, d in ((A, ea, da), (B, eb, db)):
            pk, mag = face_packaged_data(system, cl["N"], cl["u"][e], cl["invjac"][e],
                                         cl["static"][e], d)
            n = cl["N"]
            pk = pk.reshape(PK, n, n)
            if n < NM:   # project_to_mortar: interpolation in both face dimensions
                Pm = projection_matrix_parent_to_child(n, NM, MORTAR_FULL)
                pk = _apply_face_matrices(pk, Pm, Pm)
            side.append((pk.reshape(PK, NM * NM), mag))
        a2b = orient_face_map(NM, perm)
        b2a = np.argsort(a2b)
        for own, other, cl, e, d, omap, dt in ((0, 1, A, ea, da, a2b, out[ca]),
                                               (1, 0, B, eb, db, b2a, out[cb])):
            n = cl["N"]
            pk_own = _c(side[own][0])
            pk_ext = _c(side[other][0][:, omap])
            corr = np.zeros((C, NM * NM))
            if system == 0:
                L.orc_sw_boundary_terms(NM * NM, _p(pk_own), _p(pk_ext), _p(corr))
            else:
                L.orc_gh_boundary_terms(NM * NM, _p(pk_own), _p(pk_ext), _p(corr))
            corr = corr.reshape(C, NM, NM)
            if n < NM:   # project_from_mortar: L2 projection
                Rm = projection_matrix_child_to_parent(NM, n, MORTAR_FULL)
                corr = _apply_face_matrices(corr, Rm, Rm)
            lift = -0.5 * n * (n - 1) * side[own][1]      # LiftFlux.hpp:57-61
            dt[e][:, face_point_indices(n, d)] += corr.reshape(C, n * n) * lift
    return out
