"""Analytic initial data for the configurations in BASELINE.json (host side,
evaluated once; the reference evaluates the same closed forms in
Evolution/Initialization/SetVariables.hpp through its AnalyticSolutions):

  plane_wave          PointwiseFunctions/AnalyticSolutions/WaveEquation/
                      PlaneWave.cpp:56-119 with MathFunctions::Sinusoid
  gauge_wave          AnalyticSolutions/GeneralRelativity/GaugeWave.hpp:34-50
  kerr_schild         AnalyticSolutions/GeneralRelativity/KerrSchild.hpp (a = 0)
  gh_variables        GeneralizedHarmonic/{Phi.cpp:25-48, Pi.cpp:26-55}
  gaussian_plus_constant  ConstraintDamping/GaussianPlusConstant.cpp

Arrays are [..., component, point] in the reference's Variables order.
"""
from __future__ import annotations

import numpy as np

_SYM = [(a, b) for a in range(4) for b in range(a, 4)]


def plane_wave(x, t, wave_vector=(1.0, 1.0, 1.0), center=(0.0, 0.0, 0.0), amplitude=1.0,
               wavenumber=1.0, phase=0.0):
    k = np.asarray(wave_vector, float)
    omega = float(np.sqrt(k @ k))
    arg = -omega * t
    for i in range(3):
        arg = arg + k[i] * (x[..., i, :] - center[i])
    prof = amplitude * np.sin(wavenumber * arg + phase)
    dprof = amplitude * wavenumber * np.cos(wavenumber * arg + phase)
    u = np.empty(x.shape[:-2] + (5, x.shape[-1]))
    u[..., 0, :] = prof
    u[..., 1, :] = omega * dprof
    for i in range(3):
        u[..., 2 + i, :] = k[i] * dprof
    return u


def gh_variables(g, dt_g, d_g):
    """g, dt_g: dict (a,b)->array for a<=b;  d_g: dict (i,a,b)->array.
    Returns [..., 50, n]."""
    shape = next(iter(g.values())).shape
    G = np.zeros((4, 4) + shape)
    DT = np.zeros((4, 4) + shape)
    DG = np.zeros((3, 4, 4) + shape)
    for (a, b) in _SYM:
        G[a, b] = G[b, a] = g.get((a, b), 0.0)
        DT[a, b] = DT[b, a] = dt_g.get((a, b), 0.0)
        for i in range(3):
            DG[i, a, b] = DG[i, b, a] = d_g.get((i, a, b), 0.0)
    gam = np.moveaxis(G[1:, 1:], (0, 1), (-2, -1))
    inv = np.moveaxis(np.linalg.inv(gam), (-2, -1), (0, 1))
    shift = np.einsum("ij...,j...->i...", inv, G[1:, 0])
    lapse = np.sqrt(-G[0, 0] + np.einsum("i...,i...->...", shift, G[1:, 0]))
    pi = -(DT - np.einsum("i...,iab...->ab...", shift, DG)) / lapse
    u = np.zeros(shape[:-1] + (50, shape[-1]))
    for s, (a, b) in enumerate(_SYM):
        u[..., s, :] = G[a, b]
        u[..., 10 + s, :] = pi[a, b]
        for i in range(3):
            u[..., 20 + i + 3 * s, :] = DG[i, a, b]
    return u


def gauge_wave(x, t, amplitude=0.1, wavelength=1.0):
    omega = 2.0 * np.pi / wavelength
    xx = x[..., 0, :]
    H = 1.0 - amplitude * np.sin(omega * (xx - t))
    dH = -omega * amplitude * np.cos(omega * (xx - t))
    one = np.ones_like(H)
    g = {(0, 0): -H, (1, 1): H, (2, 2): one, (3, 3): one}
    dt_g = {(0, 0): dH, (1, 1): -dH}
    d_g = {(0, 0, 0): -dH, (0, 1, 1): dH}
    return gh_variables(g, dt_g, d_g)


def kerr_schild(x, mass=1.0, center=(0.0, 0.0, 0.0)):
    xc = [x[..., i, :] - center[i] for i in range(3)]
    r = np.sqrt(xc[0] ** 2 + xc[1] ** 2 + xc[2] ** 2)
    H = mass / r
    l = [np.ones_like(r)] + [xc[i] / r for i in range(3)]
    dH = [-mass * xc[i] / r ** 3 for i in range(3)]
    dl = [[np.zeros_like(r)] + [((1.0 if i == j else 0.0) - xc[i] * xc[j] / r ** 2) / r
                                 for j in range(3)] for i in range(3)]
    eta = np.diag([-1.0, 1.0, 1.0, 1.0])
    g, d_g = {}, {}
    for (a, b) in _SYM:
        g[(a, b)] = eta[a, b] + 2.0 * H * l[a] * l[b]
        for i in range(3):
            d_g[(i, a, b)] = 2.0 * dH[i] * l[a] * l[b] + 2.0 * H * (dl[i][a] * l[b]
                                                                     + l[a] * dl[i][b])
    return gh_variables(g, {}, d_g)


def superposed_kerr_schild(x, masses=(0.5, 0.5), centers=((8.0, 0.0, 0.0), (-8.0, 0.0, 0.0))):
    """Synthetic binary-black-hole-like data for BASELINE configs[4]: the Kerr-Schild
    perturbations of two non-spinning holes at rest added on the flat background,
    g = eta + sum_k 2 H_k l_k l_k, with d_t g = 0 (the reference has no GH analytic data of
    this kind; its closest relative is the superposed-Kerr-Schild free data of
    PointwiseFunctions/AnalyticData/Xcts/Binary.hpp).  Not a solution: used as initial and
    boundary data of the throughput / parity workload only."""
    eta = np.diag([-1.0, 1.0, 1.0, 1.0])
    shape = x.shape[:-2] + (x.shape[-1],)
    g = {ab: np.full(shape, eta[ab]) for ab in _SYM}
    d_g = {(i,) + ab: np.zeros(shape) for ab in _SYM for i in range(3)}
    for mass, center in zip(masses, centers):
        xc = [x[..., i, :] - center[i] for i in range(3)]
        r = np.sqrt(xc[0] ** 2 + xc[1] ** 2 + xc[2] ** 2)
        H = mass / r
        l = [np.ones_like(r)] + [xc[i] / r for i in range(3)]
        dH = [-mass * xc[i] / r ** 3 for i in range(3)]
        dl = [[np.zeros_like(r)] + [((1.0 if i == j else 0.0) - xc[i] * xc[j] / r ** 2) / r
                                     for j in range(3)] for i in range(3)]
        for (a, b) in _SYM:
            g[(a, b)] = g[(a, b)] + 2.0 * H * l[a] * l[b]
            for i in range(3):
                d_g[(i, a, b)] = d_g[(i, a, b)] + 2.0 * dH[i] * l[a] * l[b] + 2.0 * H * (
                    dl[i][a] * l[b] + l[a] * dl[i][b])
    return gh_variables(g, {}, d_g)


def gaussian_plus_constant(x, constant, amplitude, width, center=(0.0, 0.0, 0.0)):
    r2 = sum((x[..., i, :] - center[i]) ** 2 for i in range(3))
    return constant + amplitude * np.exp(-r2 / width ** 2)
