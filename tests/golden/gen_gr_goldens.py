"""Fixture generator (authoring container only; needs /root/reference): outputs of the
REFERENCE's numpy twins for the analytic data and pointwise GR/GH quantities the path
consumes, at seeded random inputs.  Writes tests/golden/gr_pointwise.npz.

  tests/Unit/PointwiseFunctions/AnalyticSolutions/GeneralRelativity/GaugeWave.py
  tests/Unit/PointwiseFunctions/GeneralRelativity/ComputeSpacetimeQuantities.py
  tests/Unit/PointwiseFunctions/GeneralRelativity/ComputeGhQuantities.py
  tests/Unit/Evolution/Systems/GeneralizedHarmonic/TestFunctions.py

Run:  python tests/golden/gen_gr_goldens.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/tests/Unit")
from Evolution.Systems.GeneralizedHarmonic import TestFunctions as ghtf  # noqa: E402
from PointwiseFunctions.AnalyticSolutions.GeneralRelativity import GaugeWave as gw  # noqa: E402
from PointwiseFunctions.GeneralRelativity import ComputeGhQuantities as ghq  # noqa: E402
from PointwiseFunctions.GeneralRelativity import ComputeSpacetimeQuantities as stq  # noqa: E402

rng = np.random.default_rng(20240929)
out = {}

# ---- (A) GaugeWave -> spacetime metric, Pi, Phi (WrappedGr.tpp:100-120 composition) ----
cases = []
for amplitude, wavelength in ((0.1, 1.0), (0.35, 1.7)):
    for _ in range(12):
        x = rng.uniform(-2.0, 2.0, 3)
        t = rng.uniform(-1.0, 3.0)
        lapse = gw.gauge_wave_lapse(x, t, amplitude, wavelength)
        dt_lapse = gw.gauge_wave_dt_lapse(x, t, amplitude, wavelength)
        d_lapse = gw.gauge_wave_d_lapse(x, t, amplitude, wavelength)
        shift = gw.gauge_wave_shift(x, t, amplitude, wavelength)
        dt_shift = gw.gauge_wave_dt_shift(x, t, amplitude, wavelength)
        d_shift = gw.gauge_wave_d_shift(x, t, amplitude, wavelength)
        gamma = gw.gauge_wave_spatial_metric(x, t, amplitude, wavelength)
        dt_gamma = gw.gauge_wave_dt_spatial_metric(x, t, amplitude, wavelength)
        d_gamma = gw.gauge_wave_d_spatial_metric(x, t, amplitude, wavelength)
        g = stq.spacetime_metric(lapse, shift, gamma)
        phi = ghq.phi(lapse, d_lapse, shift, d_shift, gamma, d_gamma)
        pi = ghq.pi(lapse, dt_lapse, shift, dt_shift, gamma, dt_gamma, phi)
        cases.append((x, t, amplitude, wavelength, g, pi, phi))
out["gw_x"] = np.array([c[0] for c in cases])
out["gw_t"] = np.array([c[1] for c in cases])
out["gw_amplitude"] = np.array([c[2] for c in cases])
out["gw_wavelength"] = np.array([c[3] for c in cases])
out["gw_spacetime_metric"] = np.array([c[4] for c in cases])
out["gw_pi"] = np.array([c[5] for c in cases])
out["gw_phi"] = np.array([c[6] for c in cases])

# ---- (B) 3+1 quantities and constraints of random physical GH states ----
n = 24
G, PI, PHI, DPHI, H = [], [], [], [], []
LAPSE, SHIFT, INVG, NVEC, NFORM, GAMMA_A, C1, C4 = [], [], [], [], [], [], [], []
for _ in range(n):
    a = rng.uniform(-0.3, 0.3, (3, 3))
    gamma = np.eye(3) + 0.5 * (a + a.T)
    lapse = rng.uniform(0.6, 1.6)
    shift = rng.uniform(-0.5, 0.5, 3)
    g = stq.spacetime_metric(lapse, shift, gamma)
    sym = lambda m: 0.5 * (m + np.swapaxes(m, -1, -2))
    pi = sym(rng.uniform(-1, 1, (4, 4)))
    phi = sym(rng.uniform(-1, 1, (3, 4, 4)))
    d_phi = sym(rng.uniform(-1, 1, (3, 3, 4, 4)))
    h = rng.uniform(-1, 1, 4)
    inv_gamma = np.linalg.inv(g[1:, 1:])
    sh = stq.shift(g, inv_gamma)
    la = stq.lapse(sh, g)
    inv_g = stq.inverse_spacetime_metric(la, sh, inv_gamma)
    nvec = stq.spacetime_normal_vector(la, sh)
    nform = stq.spacetime_normal_one_form(la, sh)
    gamma_a = ghq.trace_christoffel(nform, nvec, inv_gamma, inv_g, pi, phi)
    c1 = ghtf.gauge_constraint(h, nform, nvec, inv_gamma, inv_g, pi, phi)
    c4 = ghtf.four_index_constraint(d_phi)
    for lst, v in ((G, g), (PI, pi), (PHI, phi), (DPHI, d_phi), (H, h), (LAPSE, la), (SHIFT, sh),
                   (INVG, inv_g), (NVEC, nvec), (NFORM, nform), (GAMMA_A, gamma_a), (C1, c1),
                   (C4, c4)):
        lst.append(v)
for name, lst in (("spacetime_metric", G), ("pi", PI), ("phi", PHI), ("d_phi", DPHI),
                  ("gauge_function", H), ("lapse", LAPSE), ("shift", SHIFT),
                  ("inverse_spacetime_metric", INVG), ("spacetime_normal_vector", NVEC),
                  ("spacetime_normal_one_form", NFORM), ("trace_christoffel", GAMMA_A),
                  ("gauge_constraint", C1), ("four_index_constraint", C4)):
    out["st_" + name] = np.array(lst)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gr_pointwise.npz")
np.savez(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()})
