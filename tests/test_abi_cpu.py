"""CPU-only checks of the C-ABI library: it loads, exports every symbol that
include/dgrhs.h declares, its host-only entry points agree with the oracle, and
it fails loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dgrhs.h")).read()
    declared = set(re.findall(r"\b(dgrhs_[a-z0-9_]+)\s*\(", header))
    declared.discard("dgrhs_ctx")
    assert declared == set(lib.EXPORTS)
    handle = lib.load()
    for name in declared:
        assert getattr(handle, name) is not None


@pytest.mark.parametrize("N", [2, 3, 5, 6, 8, 12])
def test_spectral_host_functions_match_oracle(N):
    x, w = lib.collocation_points_and_weights(N)
    xo, wo = orc.lgl_points_and_weights(N)
    np.testing.assert_allclose(x, xo, rtol=0, atol=1e-15)
    np.testing.assert_allclose(w, wo, rtol=1e-14, atol=0)
    np.testing.assert_allclose(lib.differentiation_matrix(N), orc.differentiation_matrix(N),
                               rtol=1e-13, atol=1e-13)


def test_adams_bashforth_coefficients_match_oracle():
    for times, a, b in (([0.0, 1.0, 2.0], 2.0, 3.0), ([0.2, 0.4, 0.0], 0.0, 0.6),
                        ([0.4, 0.0, 0.6], 0.6, 1.2), ([0.0, 0.7, 1.0, 1.9], 1.9, 2.5)):
        got = lib.adams_bashforth_coefficients(times, a, b)
        ref = orc.ab_coefficients(times, a, b)
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-15)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.DgrhsError, match="no CPU fallback"):
        lib.Context(lib.SYSTEM_SCALAR_WAVE, 4, 8)
    with pytest.raises(lib.DgrhsError, match="no CPU fallback"):
        lib.partial_derivatives(4, np.zeros((1, 64)), np.zeros((9, 64)))


def test_product_analytic_data_matches_oracle():
    """The product-side initial data (spectre_b200.analytic) against the
    oracle's independent restatement."""
    N = 5
    brick = domain.Brick([0.5, 0.5, 0.5], [2.5] * 3, [1, 1, 1], N)
    x = brick.coords()
    for e in range(brick.n_elements):
        ref = orc.gh_vars_from_metric(*orc.gauge_wave_metric(x[e], 0.3))
        np.testing.assert_allclose(analytic.gauge_wave(x[e], 0.3), ref, rtol=1e-13, atol=1e-14)
        ref = orc.gh_vars_from_metric(*orc.kerr_schild_metric(x[e]))
        np.testing.assert_allclose(analytic.kerr_schild(x[e]), ref, rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(analytic.plane_wave(x[e], 0.2), orc.plane_wave(x[e], 0.2),
                                   rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("N", [3, 5, 8, 12])
def test_exponential_filter_matrix(N):
    """Filter matrix vs the oracle; with KerrSchild.yaml's (36, 64) the top
    Legendre mode is removed (exp(-36)) and mode N-2 is damped by ~1e-7 at most,
    so low-degree polynomials pass unchanged."""
    F = lib.exponential_filter_matrix(N, 36.0, 64)
    np.testing.assert_allclose(F, orc.exponential_filter_matrix(N, 36.0, 64), atol=1e-12)
    x, _ = lib.collocation_points_and_weights(N)
    for p in range(max(N - 2, 1)):
        np.testing.assert_allclose(F @ x ** p, x ** p, atol=1e-11)
    top = np.polynomial.legendre.legval(x, [0] * (N - 1) + [1])
    assert np.max(np.abs(F @ top)) < 1e-14


# TimeStepper::order / number_of_substeps / number_of_past_steps / stable_step of the
# reference: AdamsBashforth.cpp:60-95 (stable steps 1, 1/2, 3/11, 3/20, 45/551, 5/114
# follow from its alternating coefficient sum; orders 7 and 8: 945/40633, 945/77432), Rk3HesthavenSsp.cpp:21-26,
# Rk3Owren.cpp:8-15, Rk3Kennedy.cpp:8-10, ClassicalRungeKutta4.cpp:10-22,
# DormandPrince5.cpp:8-19
_RK3_STABLE = 0.5 * (1.0 + np.cbrt(4.0 + np.sqrt(17.0)) - 1.0 / np.cbrt(4.0 + np.sqrt(17.0)))
STEPPER_PROPERTIES = {
    "AdamsBashforth1": (lib.STEPPER_ADAMS_BASHFORTH, 1, (1, 1, 0, 1.0)),
    "AdamsBashforth2": (lib.STEPPER_ADAMS_BASHFORTH, 2, (2, 1, 1, 1.0 / 2.0)),
    "AdamsBashforth3": (lib.STEPPER_ADAMS_BASHFORTH, 3, (3, 1, 2, 3.0 / 11.0)),
    "AdamsBashforth4": (lib.STEPPER_ADAMS_BASHFORTH, 4, (4, 1, 3, 3.0 / 20.0)),
    "AdamsBashforth5": (lib.STEPPER_ADAMS_BASHFORTH, 5, (5, 1, 4, 45.0 / 551.0)),
    "AdamsBashforth6": (lib.STEPPER_ADAMS_BASHFORTH, 6, (6, 1, 5, 5.0 / 114.0)),
    "AdamsBashforth7": (lib.STEPPER_ADAMS_BASHFORTH, 7, (7, 1, 6, 945.0 / 40633.0)),
    "AdamsBashforth8": (lib.STEPPER_ADAMS_BASHFORTH, 8, (8, 1, 7, 945.0 / 77432.0)),
    "Rk3HesthavenSsp": (lib.STEPPER_RK3_HESTHAVEN, 0, (3, 3, 0, _RK3_STABLE)),
    "Rk3Owren": (lib.STEPPER_RK3_OWREN, 0, (3, 3, 0, 1.2563726633091645)),
    "Rk3Kennedy": (lib.STEPPER_RK3_KENNEDY, 0, (3, 4, 0, 1.832102281377816)),
    "ClassicalRungeKutta4": (lib.STEPPER_RK4, 0, (4, 4, 0, 1.3926467817026411)),
    "DormandPrince5": (lib.STEPPER_DORMAND_PRINCE5, 0, (5, 6, 0, 1.6532839463174733)),
}


def test_stepper_properties_match_reference_constants():
    """The library derives the stable step from the stability polynomial of its own
    tableaus / the Adams-Bashforth characteristic polynomial; the reference states
    the numbers.  Agreement pins the tableaus' stability polynomials too."""
    for name, (stepper, order, want) in STEPPER_PROPERTIES.items():
        got = lib.stepper_properties(stepper, order)
        assert got[:3] == want[:3], name
        assert abs(got[3] - want[3]) < 1e-13 * want[3], (name, got[3], want[3])
        # the oracle's tableaus have the same number of substeps
        if name in orc.RK_TABLEAUS:
            assert len(orc.RK_TABLEAUS[name][2]) == want[1]
    with pytest.raises(lib.DgrhsError, match="order must be in"):
        lib.stepper_properties(lib.STEPPER_ADAMS_BASHFORTH, 9)
    with pytest.raises(lib.DgrhsError, match="unknown time stepper"):
        lib.stepper_properties(17)


def test_cpp_time_stepper_shims_host_only():
    """TimeSteppers::AdamsBashforth / Rk3HesthavenSsp / ... of SpectreShims.hpp: the
    property queries need no GPU."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "_build", "stepper_properties")
    src = os.path.join(ROOT, "tests", "helpers", "stepper_properties.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-o", exe, src, "-L",
                           os.path.join(ROOT, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(ROOT, "spectre_b200")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = {ln.split()[0]: ln.split()[1:] for ln in out.stdout.strip().splitlines()}
    assert rows.pop("bad_order_throws") == ["1"]
    assert set(rows) == set(STEPPER_PROPERTIES)
    for name, (_, _, want) in STEPPER_PROPERTIES.items():
        o, s, p, st = rows[name]
        assert (int(o), int(s), int(p)) == want[:3]
        assert abs(float(st) - want[3]) < 1e-13 * want[3]


def _build_time_types_test():
    import subprocess
    exe = os.path.join(ROOT, "tests", "_build", "time_types_test")
    src = os.path.join(ROOT, "tests", "helpers", "time_types_test.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "spectre_b200", "host"),
                           "-o", exe, src, "-L", os.path.join(ROOT, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(ROOT, "spectre_b200")])
    return exe


def test_time_value_types_reference_known_answers():
    """Rational / Slab / Time / TimeDelta / TimeStepId of host/SpectreTime.hpp against the
    known answers of the reference's Test_Slab.cpp, Test_Time.cpp, Test_TimeStepId.cpp
    (round-off-prone slab ends included) and the id sequences of next_time_id for every
    stepper; host only."""
    import subprocess
    out = subprocess.run([_build_time_types_test()], capture_output=True, text=True)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr


def test_substep_fractions():
    import ctypes
    h = lib.load()
    for stepper, want in ((lib.STEPPER_RK3_HESTHAVEN, [1.0, 0.5]),
                          (lib.STEPPER_RK3_OWREN, [12.0 / 23.0, 4.0 / 5.0]),
                          (lib.STEPPER_RK4, [0.5, 0.5, 1.0]),
                          (lib.STEPPER_DORMAND_PRINCE5, [0.2, 0.3, 0.8, 8.0 / 9.0, 1.0])):
        buf = (ctypes.c_double * 8)()
        assert h.dgrhs_stepper_substep_fractions(stepper, buf) == 0
        assert list(buf[:len(want)]) == want
