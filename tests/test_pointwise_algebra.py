"""CPU check of the regrouped per-point algebra used inside the CUDA kernels
(spectre_b200/csrc/pointwise.cuh, compiled as host code by a test-only harness)
against the oracle, which follows the reference's operation order."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from tests.test_oracle_pins import _random_physical_gh_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(ROOT, "tests", "_build", "libharness.so")
    src = os.path.join(ROOT, "tests", "helpers", "cpu_harness.cpp")
    dep = os.path.join(ROOT, "spectre_b200", "csrc", "pointwise.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src),
                                                              os.path.getmtime(dep)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", out, src])
    return ctypes.CDLL(out)


P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def _maxrel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _du_from_logical(dlog, J, C, n):
    du = np.zeros((3 * C, n))
    for c in range(C):
        for i in range(3):
            du[3 * c + i] = (J[0 + 3 * i] * dlog[3 * c + 0] + J[1 + 3 * i] * dlog[3 * c + 1]
                             + J[2 + 3 * i] * dlog[3 * c + 2])
    return du


@pytest.mark.parametrize("gauge", [0, 1, 2])
def test_gh_volume_algebra(harness, gauge):
    rng = np.random.default_rng(11 + gauge)
    n = 64
    u = _random_physical_gh_state(rng, n)
    dlog = rng.uniform(-0.5, 0.5, (150, n))
    J = rng.uniform(-1, 1, (9, n))
    gam = rng.uniform(-1, 1, (3, n))
    H = rng.uniform(-1, 1, (4, n)); dH = rng.uniform(-1, 1, (16, n))
    dt = np.zeros((50, n))
    dhp = np.array([12.0, 1.2, 1.5, 1.7, 2, 4, 6])  # like Test_DuDt.cpp's DampedHarmonic
    coords = rng.uniform(-3, 3, (3, n))
    harness.h_gh_volume(n, gauge, P(u), P(dlog), P(J), P(gam), P(H), P(dH), P(dhp), P(coords),
                        P(dt))
    du = _du_from_logical(dlog, J, 50, n)
    gp = [orc.GAUGE_HARMONIC, orc.GAUGE_GIVEN, np.concatenate([[2.0], dhp])][gauge]
    ref = orc.gh_time_derivative(u, du, gam[0], gam[1], gam[2], gauge_params=gp, H=H, dH=dH,
                                 coords=coords)
    for blk in (slice(0, 10), slice(10, 20), slice(20, 50)):
        assert _maxrel(dt[blk], ref[blk]) < 1e-13


def _oracle_gh_face(N, ui, ue, ni, ne, gi, ge):
    f = ui.shape[1]
    L = orc.lib()
    pk = []
    mags = []
    for u, nn, gg in ((ui, ni, gi), (ue, ne, ge)):
        geo = orc.gh_geometry(u)
        ig = np.zeros((3, 3, f))
        k = 0
        for i in range(3):
            for j in range(i, 3):
                ig[i, j] = ig[j, i] = geo["inv_gamma"][k]; k += 1
        nup = np.einsum("ij...,j...->i...", ig, nn)
        mag = np.sqrt(np.einsum("i...,i...->...", nup, nn))
        nlo = np.ascontiguousarray(nn / mag); nup = np.ascontiguousarray(nup / mag)
        out = np.zeros((134, f))
        lapse = np.ascontiguousarray(geo["lapse"]); shift = np.ascontiguousarray(geo["shift"])
        g1 = np.ascontiguousarray(gg[0]); g2 = np.ascontiguousarray(gg[1])
        L.orc_gh_package_data(f, P(u), P(g1), P(g2), P(lapse), P(shift), P(nlo), P(nup), P(out))
        pk.append(out); mags.append(mag)
    corr = np.zeros((50, f))
    L.orc_gh_boundary_terms(f, P(pk[0]), P(pk[1]), P(corr))
    return corr * (-0.5 * N * (N - 1) * mags[0])


def test_gh_face_algebra(harness):
    rng = np.random.default_rng(5)
    f, N = 64, 8
    ui = _random_physical_gh_state(rng, f)
    ue = _random_physical_gh_state(rng, f)
    # make some char speeds change sign: large shifts on part of the points
    ni = rng.uniform(-1, 1, (3, f)); ne = -ni + 0.01 * rng.uniform(-1, 1, (3, f))
    gi = rng.uniform(-1.5, 1, (2, f)); ge = rng.uniform(-1.5, 1, (2, f))
    corr = np.zeros((50, f))
    harness.h_gh_face(f, N, P(ui), P(ue), P(ni), P(ne), P(gi), P(ge), P(corr))
    ref = _oracle_gh_face(N, ui, ue, ni, ne, gi, ge)
    assert _maxrel(corr, ref) < 1e-13


def test_sw_algebra(harness):
    rng = np.random.default_rng(3)
    n = 50
    u = rng.uniform(-1, 1, (5, n)); dlog = rng.uniform(-1, 1, (15, n))
    J = rng.uniform(-1, 1, (9, n)); g2 = rng.uniform(0, 1, n)
    dt = np.zeros((5, n))
    harness.h_sw_volume(n, P(u), P(dlog), P(J), P(g2), P(dt))
    ref = orc.sw_time_derivative(u, _du_from_logical(dlog, J, 5, n), g2)
    assert _maxrel(dt, ref) < 1e-13
    # face
    ui = rng.uniform(-1, 1, (5, n)); ue = rng.uniform(-1, 1, (5, n))
    ni = rng.uniform(-1, 1, (3, n)); ni /= np.linalg.norm(ni, axis=0)
    ne = np.ascontiguousarray(-ni)
    g2e = rng.uniform(0, 1, n)
    corr = np.zeros((5, n))
    harness.h_sw_face(n, P(ui), P(ue), P(ni), P(ne), P(g2), P(g2e), P(corr))
    L = orc.lib()
    pki = np.zeros((16, n)); pke = np.zeros((16, n)); ref = np.zeros((5, n))
    L.orc_sw_package_data(n, P(ui), P(g2), P(ni), P(pki))
    L.orc_sw_package_data(n, P(ue), P(g2e), P(ne), P(pke))
    L.orc_sw_boundary_terms(n, P(pki), P(pke), P(ref))
    assert _maxrel(corr, ref) < 1e-14


def test_gh_volume_split_algebra(harness):
    """The context-kernel + streaming-kernel split (inertial derivatives, context
    rebuilt from g) against the oracle."""
    rng = np.random.default_rng(21)
    n = 64
    u = _random_physical_gh_state(rng, n)
    dlog = rng.uniform(-0.5, 0.5, (150, n))
    J = rng.uniform(-1, 1, (9, n))
    gam = rng.uniform(-1, 1, (3, n))
    dt = np.zeros((50, n))
    harness.h_gh_volume_split(n, P(u), P(dlog), P(J), P(gam), P(dt))
    ref = orc.gh_time_derivative(u, _du_from_logical(dlog, J, 50, n), gam[0], gam[1], gam[2])
    for blk in (slice(0, 10), slice(10, 20), slice(20, 50)):
        assert _maxrel(dt[blk], ref[blk]) < 1e-13


@pytest.mark.parametrize("physical", [False, True])
def test_bjorhus_algebra_vs_reference_fixtures(harness, golden_dir, physical):
    """The product's pointwise ConstraintPreservingBjorhus algebra
    (spectre_b200/csrc/bjorhus.cuh, compiled for the host), both types, against
    the fixtures made from the reference's Bjorhus.py with independent random
    tensors for every argument."""
    z = np.load(os.path.join(golden_dir, "bjorhus.npz"))
    n = len(z["in_lapse"])
    c = np.ascontiguousarray
    P = lambda a: c(a, dtype=np.float64).ctypes.data_as(ctypes.c_void_p)
    out_g, out_pi, out_phi = np.zeros((n, 4, 4)), np.zeros((n, 4, 4)), np.zeros((n, 3, 4, 4))
    keys = ["normal_covector", "spacetime_metric", "pi", "phi", "coords", "gamma1", "gamma2",
            "lapse", "shift", "inverse_spacetime_metric", "spacetime_unit_normal_vector",
            "three_index_constraint", "gauge_source", "spacetime_deriv_gauge_source",
            "dt_spacetime_metric", "dt_pi", "dt_phi", "d_pi", "d_phi"]
    arrays = [c(z["in_" + k], dtype=np.float64) for k in keys]
    harness.h_bjorhus_cp(n, int(physical), *[a.ctypes.data_as(ctypes.c_void_p) for a in arrays],
                         P(out_g), P(out_pi), P(out_phi))
    want_pi = z["out_phys_corr_pi"] if physical else z["out_corr_pi"]
    want_phi = z["out_phys_corr_phi"] if physical else z["out_corr_phi"]
    scale = np.abs(want_pi).max()
    assert np.max(np.abs(out_g - z["out_corr_g"])) < 1e-12 * scale
    assert np.max(np.abs(out_pi - want_pi)) < 1e-12 * scale
    assert np.max(np.abs(out_phi - want_phi)) < 1e-12 * scale
