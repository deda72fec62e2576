"""Pin the CPU oracle against the reference's own golden vectors (CPU only)."""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc


def _sym4(a, b):
    return orc.sym4(a, b)


def _sym3(a, b):
    if a > b:
        a, b = b, a
    return a * 3 - a * (a - 1) // 2 + (b - a)


def _expand(inputs, p):
    """Dense per-point arrays from the storage-order fixture."""
    v = {k: np.array([c[p] for c in comps]) for k, comps in inputs.items()}
    g = np.zeros((4, 4)); pi = np.zeros((4, 4)); inv_g = np.zeros((4, 4))
    phi = np.zeros((3, 4, 4)); dg = np.zeros((3, 4, 4)); dpi = np.zeros((3, 4, 4))
    dphi = np.zeros((3, 3, 4, 4)); chr1 = np.zeros((4, 4, 4)); chr2 = np.zeros((4, 4, 4))
    dH = np.zeros((4, 4)); inv_gam = np.zeros((3, 3))
    for a in range(4):
        for b in range(4):
            s = _sym4(a, b)
            g[a, b] = v["psi"][s]; pi[a, b] = v["pi"][s]; inv_g[a, b] = v["inverse_psi"][s]
            dH[a, b] = v["spacetime_deriv_gauge_function"][a + 4 * b]
            for i in range(3):
                phi[i, a, b] = v["phi"][i + 3 * s]
                dg[i, a, b] = v["d_psi"][i + 3 * s]
                dpi[i, a, b] = v["d_pi"][i + 3 * s]
                for j in range(3):
                    dphi[i, j, a, b] = v["d_phi"][i + 3 * (j + 3 * s)]
            for c in range(4):
                chr1[c, a, b] = v["christoffel_first_kind"][c + 4 * s]
                chr2[c, a, b] = v["christoffel_second_kind"][c + 4 * s]
    for i in range(3):
        for j in range(3):
            inv_gam[i, j] = v["inverse_spatial_metric"][_sym3(i, j)]
    return dict(g=g, pi=pi, phi=phi, dg=dg, dpi=dpi, dphi=dphi, dH=dH, inv_gam=inv_gam,
                inv_g=inv_g, chr1=chr1, chr2=chr2, H=v["gauge_function"],
                gamma0=v["gamma0"][0], gamma1=v["gamma1"][0], gamma2=v["gamma2"][0],
                lapse=v["lapse"][0], shift=v["shift"],
                trace_chr=v["trace_christoffel_first_kind"], nform=v["normal_one_form"],
                nvec=v["normal_vector"])


def _ref_impl(d):
    L = orc.lib()
    dt_g = np.zeros((4, 4)); dt_pi = np.zeros((4, 4)); dt_phi = np.zeros((3, 4, 4))
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    arrs = {k: c(d[k]) for k in ("g", "pi", "phi", "dg", "dpi", "dphi", "H", "dH", "shift",
                                 "inv_gam", "inv_g", "trace_chr", "chr1", "chr2", "nvec",
                                 "nform")}
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    D = ctypes.c_double
    L.orc_gh_rhs_reference_impl(
        P(arrs["g"]), P(arrs["pi"]), P(arrs["phi"]), P(arrs["dg"]), P(arrs["dpi"]),
        P(arrs["dphi"]), D(d["gamma0"]), D(d["gamma1"]), D(d["gamma2"]), P(arrs["H"]),
        P(arrs["dH"]), D(d["lapse"]), P(arrs["shift"]), P(arrs["inv_gam"]), P(arrs["inv_g"]),
        P(arrs["trace_chr"]), P(arrs["chr1"]), P(arrs["chr2"]), P(arrs["nvec"]),
        P(arrs["nform"]), P(dt_g), P(dt_pi), P(dt_phi))
    return dt_g, dt_pi, dt_phi


def test_gh_rhs_reference_impl_vs_spec(golden_dir):
    """Test_DuDt.cpp:306-464: 100 SpEC numbers, reference tolerance `approx`
    (1e-14 relative-ish; one value documented at 1e-13)."""
    gold = json.load(open(os.path.join(golden_dir, "gh_dudt_spec.json")))
    out = [_ref_impl(_expand(gold["inputs"], p)) for p in range(2)]
    for e in gold["expected"]:
        dt_g, dt_pi, dt_phi = out[e["point"]]
        t = {"dt_psi": dt_g, "dt_pi": dt_pi, "dt_phi": dt_phi}[e["tensor"]]
        got = t[tuple(e["index"])]
        assert got == pytest.approx(e["value"], rel=1e-12, abs=1e-12), e


def test_upwind_penalty_vs_reference_numpy(golden_dir):
    """package_data / boundary_terms vs fixtures made by the reference's
    UpwindPenalty.py (tests/golden/gen_python_goldens.py)."""
    z = np.load(os.path.join(golden_dir, "upwind_penalty.npz"))
    L = orc.lib()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    npts = z["gh_corr"].shape[0]
    pk = []
    for side in range(2):
        u = c(z["gh_u"][side].T); g1 = c(z["gh_gamma1"][side]); g2 = c(z["gh_gamma2"][side])
        lapse = c(z["gh_lapse"][side]); shift = c(z["gh_shift"][side].T)
        nlo = c(z["gh_nlo"][side].T); nup = c(z["gh_nup"][side].T)
        out = np.zeros((134, npts))
        L.orc_gh_package_data(npts, P(u), P(g1), P(g2), P(lapse), P(shift), P(nlo), P(nup),
                              P(out))
        np.testing.assert_allclose(out.T, z["gh_packaged"][side], rtol=1e-13, atol=1e-14)
        pk.append(out)
    corr = np.zeros((50, npts))
    L.orc_gh_boundary_terms(npts, P(pk[0]), P(pk[1]), P(corr))
    np.testing.assert_allclose(corr.T, z["gh_corr"], rtol=1e-13, atol=1e-14)
    # ScalarWave
    pk = []
    for side in range(2):
        u = c(z["sw_u"][side].T); g2 = c(z["sw_gamma2"][side]); n = c(z["sw_normal"][side].T)
        out = np.zeros((16, npts))
        L.orc_sw_package_data(npts, P(u), P(g2), P(n), P(out))
        np.testing.assert_allclose(out.T, z["sw_packaged"][side], rtol=1e-13, atol=1e-14)
        pk.append(out)
    corr = np.zeros((5, npts))
    L.orc_sw_boundary_terms(npts, P(pk[0]), P(pk[1]), P(corr))
    np.testing.assert_allclose(corr.T, z["sw_corr"], rtol=1e-13, atol=1e-14)


def test_upwind_penalty_moving_mesh_vs_reference_numpy(golden_dir):
    """dg_package_data with normal_dot_mesh_velocity: GH against the fixture made by the
    reference's UpwindPenalty.py twin (gen_python_goldens.py: moving_mesh).  The ScalarWave
    twin's moving-mesh branch does not run (see the generator), so ScalarWave is pinned to
    the structure the reference's C++ states (UpwindPenalty.cpp:55-108): the speeds are
    (0, 1, -1) - n.v_g and every packaged field is its static-mesh value with the speed
    exchanged."""
    z = np.load(os.path.join(golden_dir, "upwind_penalty.npz"))
    m = np.load(os.path.join(golden_dir, "upwind_penalty_moving.npz"))
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    out = orc.gh_package_data_moving(c(z["gh_u"][0].T), z["gh_gamma1"][0], z["gh_gamma2"][0],
                                     z["gh_lapse"][0], c(z["gh_shift"][0].T),
                                     c(z["gh_nlo"][0].T), c(z["gh_nup"][0].T), m["gh_ndotv"])
    np.testing.assert_allclose(out.T, m["gh_packaged"], rtol=1e-13, atol=1e-14)
    # a zero velocity reproduces the static fixture
    out0 = orc.gh_package_data_moving(c(z["gh_u"][0].T), z["gh_gamma1"][0], z["gh_gamma2"][0],
                                      z["gh_lapse"][0], c(z["gh_shift"][0].T),
                                      c(z["gh_nlo"][0].T), c(z["gh_nup"][0].T),
                                      np.zeros_like(m["gh_ndotv"]))
    np.testing.assert_allclose(out0.T, z["gh_packaged"][0], rtol=1e-13, atol=1e-14)
    u, g2, n = c(z["sw_u"][0].T), z["sw_gamma2"][0], c(z["sw_normal"][0].T)
    nv = m["sw_ndotv"]
    pk = orc.sw_package_data_moving(u, g2, n, nv)
    st = z["sw_packaged"][0].T
    np.testing.assert_allclose(pk[13:], np.array([0.0 - nv, 1.0 - nv, -1.0 - nv]), rtol=1e-15)
    for rows, static_speed, k in (([4, 6, 7, 8], 1.0, 14), ([5, 9, 10, 11], -1.0, 15)):
        for r in rows:
            np.testing.assert_allclose(pk[r], st[r] / static_speed * pk[k], rtol=1e-13,
                                       atol=1e-14)
    ndphi = np.einsum("ip,ip->p", n, u[2:])
    np.testing.assert_allclose(pk[0], pk[13] * u[0], rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(pk[1:4], pk[13] * (u[2:] - n * ndphi), rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(pk[12], pk[13] * g2 * u[0], rtol=1e-13, atol=1e-14)


def _random_physical_gh_state(rng, n):
    """Random physical metric like TestHelpers::gr::random_lapse/shift/
    spatial_metric (Test_DuDt.cpp:489-493): lapse in (0,3), shift, SPD gamma."""
    g = np.zeros((4, 4, n))
    for p in range(n):
        lapse = rng.uniform(0.5, 2.0)
        shift = rng.uniform(-0.3, 0.3, 3)
        A = rng.uniform(-0.2, 0.2, (3, 3))
        gam = np.eye(3) + 0.5 * (A + A.T)
        shift_lo = gam @ shift
        g[0, 0, p] = -lapse ** 2 + shift @ shift_lo
        g[0, 1:, p] = g[1:, 0, p] = shift_lo
        g[1:, 1:, p] = gam
    u = np.zeros((50, n))
    for a in range(4):
        for b in range(a, 4):
            u[_sym4(a, b)] = g[a, b]
    u[10:] = rng.uniform(-0.5, 0.5, (40, n))
    return u


def test_gh_time_derivative_vs_reference_impl():
    """Same check as Test_DuDt.cpp:466-700 (eps 1e-10 there): the production
    ordering (TimeDerivative.cpp) against the SpEC-pinned reference impl on
    self-consistent geometric inputs, non-harmonic gauge with given H."""
    rng = np.random.default_rng(7)
    n = 27
    u = _random_physical_gh_state(rng, n)
    du = rng.uniform(-0.5, 0.5, (150, n))
    gam = [rng.uniform(-1, 1, n) for _ in range(3)]
    H = rng.uniform(-1, 1, (4, n)); dH = rng.uniform(-1, 1, (16, n))
    dt = orc.gh_time_derivative(u, du, *gam, gauge_params=orc.GAUGE_GIVEN, H=H, dH=dH)
    geo = orc.gh_geometry(u)
    for p in range(n):
        d = {}
        g = np.zeros((4, 4)); pi = np.zeros((4, 4)); phi = np.zeros((3, 4, 4))
        dg = np.zeros((3, 4, 4)); dpi = np.zeros((3, 4, 4)); dphi = np.zeros((3, 3, 4, 4))
        inv_g = np.zeros((4, 4)); inv_gam = np.zeros((3, 3)); dHm = np.zeros((4, 4))
        for a in range(4):
            for b in range(4):
                s = _sym4(a, b)
                g[a, b] = u[s, p]; pi[a, b] = u[10 + s, p]
                inv_g[a, b] = geo["inv_g"][s, p]
                dHm[a, b] = dH[a + 4 * b, p]
                for i in range(3):
                    phi[i, a, b] = u[20 + i + 3 * s, p]
                    dg[i, a, b] = du[3 * s + i, p]
                    dpi[i, a, b] = du[3 * (10 + s) + i, p]
                    for j in range(3):
                        dphi[i, j, a, b] = du[3 * (20 + j + 3 * s) + i, p]
        for i in range(3):
            for j in range(3):
                inv_gam[i, j] = geo["inv_gamma"][_sym3(i, j), p]
        lapse = geo["lapse"][p]; shift = geo["shift"][:, p].copy()
        dag = np.zeros((4, 4, 4))
        dag[0] = -lapse * pi + np.einsum("i,iab->ab", shift, phi)
        dag[1:] = phi
        chr1 = 0.5 * (np.einsum("ijk->kij", dag) + np.einsum("jik->kij", dag) - dag)
        chr2 = np.einsum("ad,dbc->abc", inv_g, chr1)
        trace_chr = np.einsum("abc,bc->a", chr1, inv_g)
        nvec = np.array([1.0 / lapse, *(-shift / lapse)])
        nform = np.array([-lapse, 0, 0, 0.0])
        d = dict(g=g, pi=pi, phi=phi, dg=dg, dpi=dpi, dphi=dphi, dH=dHm, inv_gam=inv_gam,
                 inv_g=inv_g, chr1=chr1, chr2=chr2, H=H[:, p].copy(), gamma0=gam[0][p],
                 gamma1=gam[1][p], gamma2=gam[2][p], lapse=lapse, shift=shift,
                 trace_chr=trace_chr, nform=nform, nvec=nvec)
        dt_g, dt_pi, dt_phi = _ref_impl(d)
        for a in range(4):
            for b in range(a, 4):
                s = _sym4(a, b)
                assert dt[s, p] == pytest.approx(dt_g[a, b], rel=1e-11, abs=1e-11)
                assert dt[10 + s, p] == pytest.approx(dt_pi[a, b], rel=1e-11, abs=1e-11)
                for i in range(3):
                    assert dt[20 + i + 3 * s, p] == pytest.approx(dt_phi[i, a, b], rel=1e-11,
                                                                  abs=1e-11)
    # harmonic gauge == given gauge with H = dH = 0
    dt_h = orc.gh_time_derivative(u, du, *gam)
    dt_0 = orc.gh_time_derivative(u, du, *gam, gauge_params=orc.GAUGE_GIVEN,
                                  H=np.zeros((4, n)), dH=np.zeros((16, n)))
    np.testing.assert_allclose(dt_h, dt_0, rtol=1e-13, atol=1e-13)


# LGL nodes and weights: tests/Unit/NumericalAlgorithms/Spectral/
# Test_LegendreGaussLobatto.cpp:33-120 (tables from Hesthaven & Warburton /
# Abramowitz & Stegun; the reference checks them at 1e-8..1e-14).
_LGL_TABLES = {
    2: ([-1.0, 1.0], [1.0, 1.0]),
    3: ([-1.0, 0.0, 1.0], [1.0 / 3.0, 4.0 / 3.0, 1.0 / 3.0]),
    4: ([-1.0, -np.sqrt(1 / 5), np.sqrt(1 / 5), 1.0], [1 / 6, 5 / 6, 5 / 6, 1 / 6]),
    5: ([-1.0, -np.sqrt(3 / 7), 0.0, np.sqrt(3 / 7), 1.0],
        [1 / 10, 49 / 90, 32 / 45, 49 / 90, 1 / 10]),
    6: ([-1.0, -np.sqrt(1 / 3 + 2 * np.sqrt(7) / 21), -np.sqrt(1 / 3 - 2 * np.sqrt(7) / 21),
         np.sqrt(1 / 3 - 2 * np.sqrt(7) / 21), np.sqrt(1 / 3 + 2 * np.sqrt(7) / 21), 1.0],
        [1 / 15, (14 - np.sqrt(7)) / 30, (14 + np.sqrt(7)) / 30, (14 + np.sqrt(7)) / 30,
         (14 - np.sqrt(7)) / 30, 1 / 15]),
}


@pytest.mark.parametrize("N", [2, 3, 4, 5, 6])
def test_lgl_tables(N):
    x, w = orc.lgl_points_and_weights(N)
    np.testing.assert_allclose(x, _LGL_TABLES[N][0], rtol=0, atol=2e-15)
    np.testing.assert_allclose(w, _LGL_TABLES[N][1], rtol=0, atol=2e-15)


@pytest.mark.parametrize("N", [4, 6, 8, 12])
def test_lgl_quadrature_and_diff_exactness(N):
    """Spectral checks of Test_Spectral.cpp: weights integrate degree 2N-3
    polynomials exactly; D differentiates degree N-1 polynomials exactly."""
    x, w = orc.lgl_points_and_weights(N)
    assert abs(w.sum() - 2.0) < 1e-14
    for p in range(0, 2 * N - 2):
        exact = 0.0 if p % 2 else 2.0 / (p + 1)
        assert abs(np.dot(w, x ** p) - exact) < 1e-13
    D = orc.differentiation_matrix(N)
    for p in range(N):
        d = D @ x ** p
        np.testing.assert_allclose(d, p * x ** max(p - 1, 0) if p else 0 * x, atol=5e-12)


def test_partial_derivatives_polynomial():
    """Test_PartialDerivatives.cpp:360-510: polynomials of degree < N are
    differentiated exactly through an affine map (tolerance 1e-10 there)."""
    N = 6
    b = orc.Brick([0.0, -1.0, 2.0], [1.0, 1.0, 5.0], [0, 0, 0], N)
    x = b.coords()[0]
    J = b.inverse_jacobian()[0]
    u = np.stack([x[0] ** 3 * x[1] ** 2 + x[2] ** 5, x[0] * x[1] * x[2], 1.0 + 0 * x[0]])
    du = orc.partial_derivatives(N, u, J)
    ex = [3 * x[0] ** 2 * x[1] ** 2, 2 * x[0] ** 3 * x[1], 5 * x[2] ** 4,
          x[1] * x[2], x[0] * x[2], x[0] * x[1], 0 * x[0], 0 * x[0], 0 * x[0]]
    np.testing.assert_allclose(du, np.stack(ex), atol=1e-9, rtol=1e-10)


def test_adams_coefficients():
    """Time/TimeSteppers/AdamsCoefficients.cpp:13-42 table vs the variable-step
    Lagrange integration (Test_AdamsCoefficients.cpp does the same comparison)."""
    for order in range(1, 7):
        times = [float(i) for i in range(order)]
        var = orc.variable_coefficients(times, times[-1], times[-1] + 1.0)
        np.testing.assert_allclose(var, orc._AB_CONST[order], rtol=1e-12, atol=1e-13)
    # non-uniform: exactness on polynomials of degree < order
    times = [0.0, 0.7, 1.0, 1.9]
    c = orc.variable_coefficients(times, 1.9, 2.5)
    for p in range(4):
        exact = (2.5 ** (p + 1) - 1.9 ** (p + 1)) / (p + 1)
        assert abs(sum(ci * t ** p for ci, t in zip(c, times)) - exact) < 1e-12


def test_damped_harmonic_vs_reference_numpy(golden_dir):
    """DampedHarmonic H_a, d_a H_b vs fixtures from the reference's
    DampedHarmonic.py (Test_DampedHarmonic.cpp compares the C++ with the same
    python at the pypp default tolerance)."""
    z = np.load(os.path.join(golden_dir, "damped_harmonic.npz"))
    L = orc.lib()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for p in range(z["g"].shape[0]):
        sigma, aL1, aL2, aS = z["params"][p]
        gp = np.array([2.0, sigma, aL1, aL2, aS, 4, 4, 4])
        H = np.zeros(4); dH = np.zeros((4, 4))
        g = np.ascontiguousarray(z["g"][p]); pi = np.ascontiguousarray(z["pi"][p])
        phi = np.ascontiguousarray(z["phi"][p]); x = np.ascontiguousarray(z["x"][p])
        L.orc_damped_harmonic(P(g), P(pi), P(phi), P(x), P(gp), P(H), P(dH))
        np.testing.assert_allclose(H, z["H"][p], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(dH, z["dH"][p], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("name,order", [("Rk3Owren", 3), ("Rk3Kennedy", 3), ("RK4", 4),
                                        ("DP5", 5), ("RK3", 3), ("AB3", 3)])
def test_stepper_convergence_order(name, order):
    """Test_{Rk3Owren,ClassicalRungeKutta4,DormandPrince5,...}.cpp check the
    convergence order with TimeStepperTestUtils: integrate y' = -y + cos t."""
    def run(dt, nsteps):
        y0 = np.array([1.0])
        ev = orc.Evolution(lambda y, t: -y + np.cos(t), y0, 0.0, dt, name)
        for _ in range(nsteps):
            ev.step()
        t = ev.time
        exact = 0.5 * (np.exp(-t) + np.cos(t) + np.sin(t))
        return abs(ev.u[0] - exact)
    e1, e2 = run(0.1, 10), run(0.05, 20)
    measured = np.log2(e1 / e2)
    assert measured > order - 0.35, (e1, e2, measured)


def test_demand_outgoing_char_speeds_vs_reference_numpy(golden_dir):
    """gh::characteristic_speeds and the DemandOutgoingCharSpeeds verdict against
    fixtures made by importing the reference's DemandOutgoingCharSpeeds.py
    (tests/golden/gen_demand_outgoing_golden.py)."""
    z = np.load(os.path.join(golden_dir, "demand_outgoing.npz"))
    for k in range(len(z["gamma1"])):
        lam = orc.gh_characteristic_speeds(z["gamma1"][k], z["lapse"][k], z["shift"][k],
                                           z["normal"][k])
        np.testing.assert_allclose(lam, z["speeds"][k], rtol=1e-14, atol=1e-15)
        assert (lam.min() < 0.0) == bool(z["violated"][k])


def test_bjorhus_constraint_preserving_vs_reference_numpy(golden_dir):
    """two_index_constraint, f_constraint and the ConstraintPreserving Bjorhus
    corrections to dt(g, Pi, Phi) against fixtures made by importing the
    reference's Bjorhus.py / TestFunctions.py with independent random tensors for
    every argument, as Test_Bjorhus.cpp feeds them
    (tests/golden/gen_bjorhus_golden.py)."""
    from oracle import bjorhus as bj
    z = np.load(os.path.join(golden_dir, "bjorhus.npz"))
    n = len(z["in_lapse"])
    incoming = 0
    for p in range(n):
        I = {k[3:]: z[k][p] for k in z.files if k.startswith("in_")}
        t_lo = np.zeros(4)
        t_lo[0] = -I["lapse"]
        ipsi, t_up = I["inverse_spacetime_metric"], I["spacetime_unit_normal_vector"]
        ig = ipsi[1:, 1:] + np.outer(I["shift"], I["shift"]) / I["lapse"] ** 2
        args = (t_lo, t_up, ig, ipsi, I["pi"], I["phi"], I["d_pi"], I["d_phi"], I["gamma2"],
                I["three_index_constraint"])
        c2 = bj.two_index_constraint(I["spacetime_deriv_gauge_source"], *args)
        np.testing.assert_allclose(c2, z["out_two_index_constraint"][p], rtol=1e-12, atol=1e-12)
        fc = bj.f_constraint(I["gauge_source"], I["spacetime_deriv_gauge_source"], *args)
        np.testing.assert_allclose(fc, z["out_f_constraint"][p], rtol=1e-12, atol=1e-12)
        speeds = bj.characteristic_speeds(I["gamma1"], I["lapse"], I["shift"],
                                          I["normal_covector"])
        np.testing.assert_allclose(speeds, z["out_char_speeds"][p], rtol=1e-14, atol=1e-15)
        incoming += speeds.min() < 0
        cg, cp, cph = bj.bjorhus_constraint_preserving(
            I["normal_covector"], I["spacetime_metric"], I["pi"], I["phi"], I["coords"],
            I["gamma1"], I["gamma2"], I["lapse"], I["shift"], ipsi, t_up,
            I["three_index_constraint"], I["gauge_source"], I["spacetime_deriv_gauge_source"],
            I["dt_spacetime_metric"], I["dt_pi"], I["dt_phi"], I["d_pi"], I["d_phi"])
        np.testing.assert_allclose(cg, z["out_corr_g"][p], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(cp, z["out_corr_pi"][p], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(cph, z["out_corr_phi"][p], rtol=1e-12, atol=1e-12)
    assert 0 < incoming < n + 1


def test_bjorhus_physical_vs_reference_numpy(golden_dir):
    """Type ConstraintPreservingPhysical (Weyl propagating mode U^{3-}, spatial
    Ricci tensor and covariant derivative of the extrinsic curvature from the GH
    variables) against the reference's Bjorhus.py."""
    from oracle import bjorhus as bj
    z = np.load(os.path.join(golden_dir, "bjorhus.npz"))
    for p in range(len(z["in_lapse"])):
        I = {k[3:]: z[k][p] for k in z.files if k.startswith("in_")}
        cg, cp, cph = bj.bjorhus_constraint_preserving(
            I["normal_covector"], I["spacetime_metric"], I["pi"], I["phi"], I["coords"],
            I["gamma1"], I["gamma2"], I["lapse"], I["shift"], I["inverse_spacetime_metric"],
            I["spacetime_unit_normal_vector"], I["three_index_constraint"], I["gauge_source"],
            I["spacetime_deriv_gauge_source"], I["dt_spacetime_metric"], I["dt_pi"], I["dt_phi"],
            I["d_pi"], I["d_phi"], physical=True)
        np.testing.assert_allclose(cg, z["out_corr_g"][p], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(cp, z["out_phys_corr_pi"][p], rtol=1e-12, atol=2e-12)
        np.testing.assert_allclose(cph, z["out_phys_corr_phi"][p], rtol=1e-12, atol=2e-12)


def _pack_gh(g, pi, phi):
    """[n,4,4], [n,4,4], [n,3,4,4] -> Variables layout [50, n]."""
    n = len(g)
    u = np.zeros((50, n))
    for a in range(4):
        for b in range(a, 4):
            s = orc.sym4(a, b)
            u[s], u[10 + s] = g[:, a, b], pi[:, a, b]
            for i in range(3):
                u[20 + i + 3 * s] = phi[:, i, a, b]
    return u


def test_gauge_wave_gh_variables_vs_reference_numpy(golden_dir):
    """The analytic GaugeWave solution composed to GH variables (spacetime metric,
    Pi, Phi as WrappedGr.tpp:100-120 does) against the reference's GaugeWave.py +
    ComputeSpacetimeQuantities.py + ComputeGhQuantities.py (tests/golden/gen_gr_goldens.py)."""
    z = np.load(os.path.join(golden_dir, "gr_pointwise.npz"))
    for k in range(len(z["gw_t"])):
        x = z["gw_x"][k][:, None]
        u = orc.gh_vars_from_metric(*orc.gauge_wave_metric(
            x, float(z["gw_t"][k]), float(z["gw_amplitude"][k]), float(z["gw_wavelength"][k])))
        want = _pack_gh(z["gw_spacetime_metric"][k:k + 1], z["gw_pi"][k:k + 1], z["gw_phi"][k:k + 1])
        np.testing.assert_allclose(u, want, rtol=1e-13, atol=1e-14)


def test_spacetime_quantities_and_constraints_vs_reference_numpy(golden_dir):
    """Lapse, shift, inverse spacetime metric, trace of the Christoffel symbols
    (-H_a of the AnalyticChristoffel gauge), gauge constraint and four-index constraint
    of random physical GH states against ComputeSpacetimeQuantities.py,
    ComputeGhQuantities.py::trace_christoffel and GeneralizedHarmonic/TestFunctions.py
    (gauge_constraint :9-52, four_index_constraint :295-300)."""
    z = {k[3:]: v for k, v in np.load(os.path.join(golden_dir, "gr_pointwise.npz")).items()
         if k.startswith("st_")}
    n = len(z["lapse"])
    u = _pack_gh(z["spacetime_metric"], z["pi"], z["phi"])
    geo = orc.gh_geometry(u)
    np.testing.assert_allclose(geo["lapse"], z["lapse"], rtol=1e-13)
    np.testing.assert_allclose(geo["shift"], z["shift"].T, rtol=1e-12, atol=1e-14)
    k = 0
    for a in range(4):
        for b in range(a, 4):
            np.testing.assert_allclose(geo["inv_g"][k], z["inverse_spacetime_metric"][:, a, b],
                                       rtol=1e-12, atol=1e-13)
            k += 1
    # H_a = -Gamma_a; the numerical derivative part needs a mesh: 8 points = one N = 2
    # element per chunk with an identity Jacobian (only H is compared)
    assert n % 8 == 0
    eye = np.zeros((9, 8))
    eye[0] = eye[4] = eye[8] = 1.0
    for c in range(n // 8):
        sl = slice(8 * c, 8 * c + 8)
        Hg, _ = orc.analytic_christoffel_gauge(2, np.ascontiguousarray(u[:, sl]), eye)
        np.testing.assert_allclose(-Hg, z["trace_christoffel"][sl].T, rtol=1e-12, atol=1e-12)
        # gauge constraint C_a = H_a + Gamma_a for the fixture's random H_a
        np.testing.assert_allclose(z["gauge_function"][sl].T - Hg, z["gauge_constraint"][sl].T,
                                   rtol=1e-12, atol=1e-12)
    # four-index constraint C_iab = eps_ijk d_j Phi_kab, the function gh_constraint_norms uses
    c4 = orc.four_index_constraint(np.moveaxis(z["d_phi"], 0, -1))      # [j, k, a, b, n]
    np.testing.assert_allclose(np.moveaxis(c4, -1, 0), z["four_index_constraint"],
                               rtol=1e-13, atol=1e-14)


def test_oracle_moving_mesh_terms_are_the_grid_frame_derivative():
    """On a moving mesh the right-hand side is the time derivative at a moving grid point:
    d_t u + v_g.grad u (VolumeTermsImpl.tpp:155-235).  For the analytic ScalarWave plane wave
    and a smooth periodic velocity the oracle reproduces it to truncation error, for Psi, Pi
    and Phi; with the static-mesh right-hand side the difference is O(|v|)."""
    from spectre_b200 import analytic, domain
    N, L = 9, 2 * np.pi
    brick = domain.Brick([0, 0, 0], [L] * 3, [1, 1, 1], N)
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    stat = np.zeros((brick.n_elements, 1, brick.n))
    u = analytic.plane_wave(x, 0.2)
    v = np.empty((brick.n_elements, 3, brick.n))
    v[:, 0] = 0.3 + 0.2 * np.sin(x[:, 1])
    v[:, 1] = -0.25 + 0.1 * np.cos(x[:, 2])
    v[:, 2] = 0.15 * np.sin(x[:, 0])
    eps = 1e-6
    d_t = (analytic.plane_wave(x, 0.2 + eps) - analytic.plane_wave(x, 0.2 - eps)) / (2 * eps)
    grad = np.zeros((3,) + u.shape)
    for i in range(3):
        xp, xm = x.copy(), x.copy()
        xp[:, i] += eps
        xm[:, i] -= eps
        grad[i] = (analytic.plane_wave(xp, 0.2) - analytic.plane_wave(xm, 0.2)) / (2 * eps)
    expected = d_t + np.einsum("eip,iecp->ecp", v, grad)
    got = orc.dg_rhs(0, N, u, J, stat, nb, mesh_velocity=v)
    static = orc.dg_rhs(0, N, u, J, stat, nb)
    assert np.max(np.abs(got - expected)) < 2e-4
    assert np.max(np.abs(static - expected)) > 0.1
