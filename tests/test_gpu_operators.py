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
