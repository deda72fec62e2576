"""Operator-level GPU parity, shaped like the reference's own unit tests:
Test_DuDt.cpp (gh::TimeDerivative vs the SpEC-pinned reference impl),
Test_UpwindPenalty.cpp (vs UpwindPenalty.py fixtures + conservation),
Test_LiftFlux.cpp, Test_TimeDerivative.cpp (ScalarWave)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import lib
from tests.test_oracle_pins import _random_physical_gh_state

pytestmark = pytest.mark.gpu


def _maxrel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("harmonic", [True, False])
def test_gh_time_derivative_apply(harmonic):
    """Test_DuDt.cpp:466-700: random physical metrics, random derivatives."""
    rng = np.random.default_rng(42)
    n = 27  # 3-point LGL mesh in 3-D like the reference test
    u = _random_physical_gh_state(rng, n)
    du = rng.uniform(-0.5, 0.5, (150, n))
    gam = rng.uniform(-1, 1, (3, n))
    H = rng.uniform(-1, 1, (4, n)); dH = rng.uniform(-1, 1, (16, n))
    if harmonic:
        got = lib.gh_time_derivative(u, du, *gam)
        ref = orc.gh_time_derivative(u, du, *gam)
    else:
        got = lib.gh_time_derivative(u, du, *gam, gauge_h=H, d4_gauge_h=dH)
        ref = orc.gh_time_derivative(u, du, *gam, gauge_params=orc.GAUGE_GIVEN, H=H, dH=dH)
    for blk in (slice(0, 10), slice(10, 20), slice(20, 50)):
        assert _maxrel(got[blk], ref[blk]) < 1e-12


def test_sw_time_derivative_apply():
    rng = np.random.default_rng(1)
    n = 125
    u = rng.uniform(-1, 1, (5, n)); du = rng.uniform(-1, 1, (15, n)); g2 = rng.uniform(0, 1, n)
    got = lib.sw_time_derivative(u, du, g2)
    assert _maxrel(got, orc.sw_time_derivative(u, du, g2)) < 1e-14


def test_upwind_penalty_vs_reference_numpy_fixtures(golden_dir):
    """The GPU operators against outputs of the reference's UpwindPenalty.py."""
    z = np.load(os.path.join(golden_dir, "upwind_penalty.npz"))
    pk = []
    for side in range(2):
        out, _ = lib.gh_package_data(z["gh_u"][side].T, z["gh_gamma1"][side],
                                     z["gh_gamma2"][side], z["gh_lapse"][side],
                                     z["gh_shift"][side].T, z["gh_nlo"][side].T,
                                     z["gh_nup"][side].T)
        np.testing.assert_allclose(out.T, z["gh_packaged"][side], rtol=1e-13, atol=1e-14)
        pk.append(out)
    np.testing.assert_allclose(lib.gh_boundary_terms(pk[0], pk[1]).T, z["gh_corr"],
                               rtol=1e-13, atol=1e-14)
    pk = []
    for side in range(2):
        out, _ = lib.sw_package_data(z["sw_u"][side].T, z["sw_gamma2"][side],
                                     z["sw_normal"][side].T)
        np.testing.assert_allclose(out.T, z["sw_packaged"][side], rtol=1e-13, atol=1e-14)
        pk.append(out)
    np.testing.assert_allclose(lib.sw_boundary_terms(pk[0], pk[1]).T, z["sw_corr"],
                               rtol=1e-13, atol=1e-14)


def test_upwind_penalty_zero_on_smooth_solution():
    """Helpers/Evolution/DiscontinuousGalerkin/BoundaryCorrections.hpp:432-480
    (ZeroOnSmoothSolution::Yes): if the solution is the same on both sides and
    the exterior normals are minus the interior ones, the StrongInertial
    correction is identically zero."""
    rng = np.random.default_rng(8)
    f = 40
    u = _random_physical_gh_state(rng, f)
    nlo = rng.uniform(-1, 1, (3, f)); nup = rng.uniform(-1, 1, (3, f))
    lapse = rng.uniform(0.5, 2, f); shift = rng.uniform(-1.5, 1.5, (3, f))
    g1 = rng.uniform(-1, 1, f); g2 = rng.uniform(-1, 1, f)
    pk_a, ms = lib.gh_package_data(u, g1, g2, lapse, shift, nlo, nup)
    pk_b, _ = lib.gh_package_data(u, g1, g2, lapse, shift, -nlo, -nup)
    corr = lib.gh_boundary_terms(pk_a, pk_b)
    assert np.max(np.abs(corr)) < 1e-12 * np.max(np.abs(pk_a))
    # the return value of dg_package_data is the largest characteristic speed
    assert ms == np.max(pk_a[130:134])
    # ScalarWave
    us = rng.uniform(-1, 1, (5, f)); n = rng.uniform(-1, 1, (3, f)); n /= np.linalg.norm(n, axis=0)
    g2s = rng.uniform(0, 1, f)
    pa, _ = lib.sw_package_data(us, g2s, n)
    pb, _ = lib.sw_package_data(us, g2s, -n)
    assert np.max(np.abs(lib.sw_boundary_terms(pa, pb))) < 1e-14


def test_lift_flux():
    """Test_LiftFlux.cpp: -0.5 N (N-1) |n| scaling on the boundary slice."""
    rng = np.random.default_rng(3)
    f, ncomp, extent = 25, 5, 5
    c = rng.uniform(-1, 1, (ncomp, f)); mag = rng.uniform(0.5, 2, f)
    got = lib.lift_flux(c, extent, mag)
    np.testing.assert_array_equal(got, c * (-0.5 * extent * (extent - 1) * mag))


def test_cpp_shims_on_gpu():
    """spectre_b200/host/SpectreShims.hpp (the C++ mirror of the reference's
    operator classes) compiled with g++ -std=c++20 against libdgrhs.so."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "shim_test")
    src = os.path.join(root, "tests", "helpers", "shim_test.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-o", exe, src, "-L",
                           os.path.join(root, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(root, "spectre_b200")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "SHIM OK" in out.stdout, out.stdout + out.stderr


def test_cpp_level2_example_matches_oracle():
    """spectre_b200/host/evolve_scalar_wave.cpp: the PlaneWave3D.yaml configuration
    built and driven entirely from C++20 through the C-ABI; its ObserveNorms-style
    errors equal those of the oracle evolution of the same configuration."""
    import subprocess
    from spectre_b200 import analytic, domain
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "evolve_scalar_wave")
    src = os.path.join(root, "spectre_b200", "host", "evolve_scalar_wave.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-o", exe, src, "-L",
                           os.path.join(root, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(root, "spectre_b200")])
    steps = 10
    out = subprocess.run([exe, str(steps)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    got = {ln.split()[0]: float(ln.split()[1]) for ln in out.stdout.strip().splitlines()}
    N, dt = 5, 1e-3
    brick = domain.Brick([0, 0, 0], [2 * np.pi] * 3, [1, 1, 1], N, order="lexicographic")
    x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
    stat = np.zeros((brick.n_elements, 1, N ** 3))
    ev = orc.Evolution(lambda v, t: orc.dg_rhs(0, N, v, J, stat, nb), analytic.plane_wave(x, 0.0),
                       0.0, dt, "AB3")
    for _ in range(steps):
        ev.step()
    assert got["time"] == pytest.approx(ev.time, abs=1e-15)
    exact = analytic.plane_wave(x, ev.time)
    npts = exact.shape[0] * exact.shape[2]
    for name, (a, b) in zip(("Psi", "Pi", "Phi"), ((0, 1), (1, 2), (2, 5))):
        want = np.sqrt(np.sum((ev.u[:, a:b] - exact[:, a:b]) ** 2) / npts)
        assert got[f"Error({name})"] == pytest.approx(want, rel=1e-8)


@pytest.mark.parametrize("physical", [False, True])
def test_bjorhus_dg_time_derivative_vs_reference_numpy_fixtures(golden_dir, physical):
    """ConstraintPreservingBjorhus::dg_time_derivative through the operator-level
    C-ABI entry point, fed like Test_Bjorhus.cpp feeds the C++ (an independent random
    tensor for every argument), against the outputs of the reference's Bjorhus.py."""
    import ctypes
    z = np.load(os.path.join(golden_dir, "bjorhus.npz"))
    n = len(z["in_lapse"])
    pairs = [(a, b) for a in range(4) for b in range(a, 4)]

    def aa(t):      # [n,4,4] -> [10][n]
        return np.ascontiguousarray(np.stack([t[:, a, b] for a, b in pairs]))

    def iaa(t):     # [n,3,4,4] -> [30][n] at i + 3 sym
        out = np.zeros((30, n))
        for s, (a, b) in enumerate(pairs):
            for i in range(3):
                out[i + 3 * s] = t[:, i, a, b]
        return out

    def ijaa(t):    # [n,3,3,4,4] -> [90][n] at i + 3 (j + 3 sym)
        out = np.zeros((90, n))
        for s, (a, b) in enumerate(pairs):
            for j in range(3):
                for i in range(3):
                    out[i + 3 * (j + 3 * s)] = t[:, i, j, a, b]
        return out

    def vec(t):     # [n,k] -> [k][n]
        return np.ascontiguousarray(t.T)
    I = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    dH = np.zeros((16, n))
    for a in range(4):
        for b in range(4):
            dH[a + 4 * b] = I["spacetime_deriv_gauge_source"][:, a, b]
    args = [vec(I["normal_covector"]), aa(I["spacetime_metric"]), aa(I["pi"]), iaa(I["phi"]),
            vec(I["coords"]), I["gamma1"].copy(), I["gamma2"].copy(), I["lapse"].copy(),
            vec(I["shift"]), aa(I["inverse_spacetime_metric"]),
            vec(I["spacetime_unit_normal_vector"]), iaa(I["three_index_constraint"]),
            vec(I["gauge_source"]), dH, aa(I["dt_spacetime_metric"]), aa(I["dt_pi"]),
            iaa(I["dt_phi"]), iaa(I["d_pi"]), ijaa(I["d_phi"])]
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in args]
    og, op, oph = np.zeros((10, n)), np.zeros((10, n)), np.zeros((30, n))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib._check(lib.load().dgrhs_gh_bjorhus_dg_time_derivative(
        n, int(physical), *[P(a) for a in args], P(og), P(op), P(oph)))
    want_pi = z["out_phys_corr_pi"] if physical else z["out_corr_pi"]
    want_phi = z["out_phys_corr_phi"] if physical else z["out_corr_phi"]
    scale = np.abs(want_pi).max()
    assert np.max(np.abs(og - aa(z["out_corr_g"]))) < 1e-12 * scale
    assert np.max(np.abs(op - aa(want_pi))) < 1e-12 * scale
    assert np.max(np.abs(oph - iaa(want_phi))) < 1e-12 * scale


@pytest.mark.parametrize("N", range(2, 13))
@pytest.mark.parametrize("system", [lib.SYSTEM_SCALAR_WAVE, lib.SYSTEM_GH])
def test_apply_exponential_filter(system, N):
    """Filters::Exponential applied once (ExponentialFilter.cpp:45-76): the line kernel
    (one thread per grid line, matrix entries from the constant bank) vs apply_matrices of
    the oracle, every N, with an element count that leaves the last CTA partly filled."""
    C = 5 if system == lib.SYSTEM_SCALAR_WAVE else 50
    E = 3
    rng = np.random.default_rng(100 * N + C)
    u = rng.uniform(-1, 1, (E, C, N ** 3))
    ctx = lib.Context(system, N, E)
    with pytest.raises(lib.DgrhsError, match="no exponential filter set"):
        ctx.apply_exponential_filter()
    for alpha, half_power in ((36.0, 64), (4.0, 2)):
        ctx.set_exponential_filter(True, alpha, half_power)
        ctx.set_state(u)
        ctx.apply_exponential_filter()
        got = ctx.get_state()
        F = orc.exponential_filter_matrix(N, alpha, half_power)
        want = orc.apply_filter(N, u, F)
        assert np.max(np.abs(got - want)) < 1e-13 * max(1.0, np.max(np.abs(want)))
    ctx.close()


def _call(name, *args):
    h = lib.load()
    rc = getattr(h, name)(*args)
    if rc:
        raise lib.DgrhsError(h.dgrhs_last_error().decode())


def _ints(v):
    import ctypes
    return (ctypes.c_int * len(v))(*v)


def _dptr(a):
    import ctypes
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


@pytest.mark.parametrize("face,mortar,sizes", [
    ((5, 5), (5, 5), (orc.MORTAR_UPPER_HALF, orc.MORTAR_FULL)),      # h-mortar in a
    ((4, 6), (4, 6), (orc.MORTAR_LOWER_HALF, orc.MORTAR_UPPER_HALF)),
    ((4, 5), (6, 5), (orc.MORTAR_FULL, orc.MORTAR_FULL)),            # p-mortar in a
    ((3, 4), (5, 7), (orc.MORTAR_FULL, orc.MORTAR_LOWER_HALF)),      # hp
    ((12, 2), (12, 12), (orc.MORTAR_UPPER_HALF, orc.MORTAR_UPPER_HALF))])
def test_project_to_and_from_mortar(face, mortar, sizes):
    """dg::project_to_mortar / project_from_mortar (MortarHelpers.hpp:74-129) through the
    C-ABI vs apply_matrices with the oracle's projection matrices (pinned to the closed
    forms of Test_Projection.cpp in tests/test_oracle_pins.py); projecting a polynomial of
    the face degree to the mortar and back gives it back when the mortar covers the face."""
    C = 3
    rng = np.random.default_rng(sum(face) + 7 * sum(mortar))
    v = rng.uniform(-1, 1, (C, face[1], face[0]))

    def mats(to_mortar):
        out = []
        for d in range(2):
            if face[d] == mortar[d] and sizes[d] == orc.MORTAR_FULL:
                out.append(None)
            elif to_mortar:
                out.append(orc.projection_matrix_parent_to_child(face[d], mortar[d], sizes[d]))
            else:
                out.append(orc.projection_matrix_child_to_parent(mortar[d], face[d], sizes[d]))
        return out

    def apply(x, ms):          # x [C][b][a]
        if ms[0] is not None:
            x = np.einsum("ta,cba->cbt", ms[0], x)
        if ms[1] is not None:
            x = np.einsum("tb,cba->cta", ms[1], x)
        return x

    want = apply(v, mats(True))
    got = np.zeros((C, mortar[1], mortar[0]))
    _call("dgrhs_project_to_mortar", C, _ints(face), _ints(mortar), _ints(sizes), _dptr(v), _dptr(got))
    assert np.max(np.abs(got - want)) < 1e-13
    back_want = apply(want, mats(False))
    back = np.zeros((C, face[1], face[0]))
    _call("dgrhs_project_from_mortar", C, _ints(face), _ints(mortar), _ints(sizes), _dptr(got), _dptr(back))
    assert np.max(np.abs(back - back_want)) < 1e-12
    if sizes == (orc.MORTAR_FULL, orc.MORTAR_FULL):
        assert np.max(np.abs(back - v)) < 1e-12     # p-mortar: the L2 projection inverts the interpolation
    with pytest.raises(lib.DgrhsError, match="mortar extent"):
        _call("dgrhs_project_to_mortar", C, _ints((6, 6)), _ints((5, 6)), _ints(sizes), _dptr(v), _dptr(got))
    with pytest.raises(lib.DgrhsError, match="no projection is needed"):
        _call("dgrhs_project_from_mortar", C, _ints((4, 4)), _ints((4, 4)), _ints((0, 0)), _dptr(v), _dptr(got))


@pytest.mark.parametrize("perm", range(8))
def test_orient_variables_on_slice(perm):
    """orient_variables_on_slice (OrientationMapHelpers.cpp:25-120) on a face with different
    extents in its two dimensions: a point (qa, qb) of this element's face lands where the
    neighbour's frame has it (the same rule the face kernels gather with)."""
    C, na, nb = 2, 3, 5
    rng = np.random.default_rng(perm)
    v = rng.uniform(-1, 1, (C, nb, na))
    got = np.zeros(C * na * nb)
    _call("dgrhs_orient_variables_on_slice", C, _ints((na, nb)), perm, _dptr(v), _dptr(got))
    ma, mb = (nb, na) if perm & 1 else (na, nb)
    want = np.zeros((C, mb, ma))
    for qb in range(nb):
        for qa in range(na):
            ta, tb = (qb, qa) if perm & 1 else (qa, qb)
            if perm & 2:
                ta = ma - 1 - ta
            if perm & 4:
                tb = mb - 1 - tb
            want[:, tb, ta] = v[:, qb, qa]
    assert np.array_equal(got.reshape(C, mb, ma), want)


def test_update_u_operator():
    """TimeStepper::update_u on a flat span (TimeStepper.hpp:96-102): Adams-Bashforth with
    unequal history spacing against the oracle's coefficients, Rk3HesthavenSsp and a Butcher
    tableau method against the oracle's substep formulas."""
    import ctypes
    from fractions import Fraction
    size = 1000
    rng = np.random.default_rng(3)
    u = rng.uniform(-1, 1, size)
    # AB3, history at t = 0, 0.3, 0.5 (dt units), step 0.5 -> 0.9
    times = np.array([0.0, 0.3, 0.5])
    f = rng.uniform(-1, 1, (3, size))
    coef = orc.ab_coefficients_frac([Fraction(0), Fraction(3, 10), Fraction(1, 2)], Fraction(1, 2),
                                    Fraction(9, 10), 1.0)
    want = u.copy()
    for c, d in zip(coef, f):
        want += c * d
    got = u.copy()
    _call("dgrhs_update_u", lib.STEPPER_ADAMS_BASHFORTH, 3, ctypes.c_longlong(size), _dptr(got), 3,
          _dptr(times), _dptr(f), None, ctypes.c_double(0.4))
    assert np.max(np.abs(got - want)) < 1e-13
    # Rk3HesthavenSsp: three substeps
    dt = 0.01
    u0 = u.copy()
    cur = u.copy()
    fs = rng.uniform(-1, 1, (3, size))
    wants = [u0 + dt * fs[0]]
    wants.append(0.25 * (3.0 * u0 + wants[0] + dt * fs[1]))
    wants.append((1.0 / 3.0) * (u0 + 2.0 * wants[1] + 2.0 * dt * fs[2]))
    for k in range(3):
        _call("dgrhs_update_u", lib.STEPPER_RK3_HESTHAVEN, 0, ctypes.c_longlong(size), _dptr(cur), k + 1,
              None, _dptr(np.ascontiguousarray(fs[:k + 1])), _dptr(u0), ctypes.c_double(dt))
        assert np.max(np.abs(cur - wants[k])) < 1e-14
    # DormandPrince5 through the tableau rows
    c_, A, b = orc.RK_TABLEAUS["DP5"]
    nsub = len(b)
    fs = rng.uniform(-1, 1, (nsub, size))
    cur = u.copy()
    for k in range(nsub):
        row = b if k == nsub - 1 else A[k]
        want = u0.copy()
        for cf, d in zip(row, fs):
            if cf != 0.0:
                want += cf * dt * d
        _call("dgrhs_update_u", lib.STEPPER_DORMAND_PRINCE5, 0, ctypes.c_longlong(size), _dptr(cur), k + 1,
              None, _dptr(np.ascontiguousarray(fs[:k + 1])), _dptr(u0), ctypes.c_double(dt))
        assert np.max(np.abs(cur - want)) < 1e-14
    with pytest.raises(lib.DgrhsError, match="needs k history entries"):
        _call("dgrhs_update_u", lib.STEPPER_ADAMS_BASHFORTH, 4, ctypes.c_longlong(size), _dptr(got), 3,
              _dptr(times), _dptr(f), None, ctypes.c_double(0.4))
