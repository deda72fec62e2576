"""Generate golden input/output vectors by importing the REFERENCE's own numpy
oracles (authoring container only; needs /root/reference):

    python tests/golden/gen_python_goldens.py

Writes tests/golden/upwind_penalty.npz: random per-point inputs and the outputs
of tests/Unit/Evolution/Systems/{GeneralizedHarmonic,ScalarWave}/
BoundaryCorrections/UpwindPenalty.py (dg_package_data, dg_boundary_terms).
Seeds are fixed; the file travels to the GPU box, the reference does not.
"""
import importlib.util
import os

import numpy as np

REF = "/root/reference/tests/Unit/Evolution/Systems"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def sym4(a, b):
    if a > b:
        a, b = b, a
    return a * 4 - a * (a - 1) // 2 + (b - a)


def pack_aa(t):  # [4,4] -> 10
    return np.array([t[a, b] for a in range(4) for b in range(a, 4)])


def pack_iaa(t):  # [3,4,4] -> 30 (i + 3*sym)
    out = np.zeros(30)
    for a in range(4):
        for b in range(a, 4):
            for i in range(3):
                out[i + 3 * sym4(a, b)] = t[i, a, b]
    return out


def main():
    gh = _load(f"{REF}/GeneralizedHarmonic/BoundaryCorrections/UpwindPenalty.py", "gh_up")
    sw = _load(f"{REF}/ScalarWave/BoundaryCorrections/UpwindPenalty.py", "sw_up")
    rng = np.random.default_rng(20240929)
    npts = 64
    out = {}
    # ---- GH ----
    gh_u = np.zeros((2, npts, 50)); gh_g1 = np.zeros((2, npts)); gh_g2 = np.zeros((2, npts))
    gh_lapse = np.zeros((2, npts)); gh_shift = np.zeros((2, npts, 3))
    gh_nlo = np.zeros((2, npts, 3)); gh_nup = np.zeros((2, npts, 3))
    gh_pk = np.zeros((2, npts, 134)); gh_corr = np.zeros((npts, 50))
    for p in range(npts):
        packs = []
        for side in range(2):
            def symm(x):
                return 0.5 * (x + x.T)
            g = symm(rng.uniform(-1, 1, (4, 4)))
            pi = symm(rng.uniform(-1, 1, (4, 4)))
            phi = rng.uniform(-1, 1, (3, 4, 4))
            phi = 0.5 * (phi + phi.transpose(0, 2, 1))
            g1, g2 = rng.uniform(-1, 1), rng.uniform(-1, 1)
            lapse = rng.uniform(0.2, 2.0)
            shift = rng.uniform(-1.5, 1.5, 3)
            nlo = rng.uniform(-1, 1, 3)
            nup = rng.uniform(-1, 1, 3)
            r = gh.dg_package_data(g, pi, phi, g1, g2, lapse, shift, nlo, nup, None, None)
            pk = np.concatenate([pack_aa(r[0]), pack_iaa(r[1]), pack_aa(r[2]), pack_aa(r[3]),
                                 pack_iaa(r[4]), pack_iaa(r[5]), pack_aa(r[6]), r[7]])
            packs.append(r)
            gh_u[side, p] = np.concatenate([pack_aa(g), pack_aa(pi), pack_iaa(phi)])
            gh_g1[side, p], gh_g2[side, p] = g1, g2
            gh_lapse[side, p] = lapse
            gh_shift[side, p] = shift
            gh_nlo[side, p], gh_nup[side, p] = nlo, nup
            gh_pk[side, p] = pk
        c = gh.dg_boundary_terms(*packs[0], *packs[1], True)
        gh_corr[p] = np.concatenate([pack_aa(c[0]), pack_aa(c[1]), pack_iaa(c[2])])
    out.update(gh_u=gh_u, gh_gamma1=gh_g1, gh_gamma2=gh_g2, gh_lapse=gh_lapse,
               gh_shift=gh_shift, gh_nlo=gh_nlo, gh_nup=gh_nup, gh_packaged=gh_pk,
               gh_corr=gh_corr)
    # ---- SW ----
    sw_u = np.zeros((2, npts, 5)); sw_g2 = np.zeros((2, npts)); sw_n = np.zeros((2, npts, 3))
    sw_pk = np.zeros((2, npts, 16)); sw_corr = np.zeros((npts, 5))
    for p in range(npts):
        packs = []
        for side in range(2):
            psi, pi = rng.uniform(-1, 1), rng.uniform(-1, 1)
            phi = rng.uniform(-1, 1, 3)
            g2 = rng.uniform(0, 1)
            n = rng.uniform(-1, 1, 3)
            n /= np.linalg.norm(n)
            r = sw.dg_package_data(psi, pi, phi, g2, n, None, None)
            packs.append(r)
            sw_u[side, p] = np.concatenate([[psi, pi], phi])
            sw_g2[side, p] = g2
            sw_n[side, p] = n
            sw_pk[side, p] = np.concatenate([[r[0]], r[1], [r[2]], [r[3]], r[4], r[5], [r[6]],
                                             r[7]])
        c = sw.dg_boundary_terms(*packs[0], *packs[1], True)
        sw_corr[p] = np.concatenate([[c[0]], [c[1]], c[2]])
    out.update(sw_u=sw_u, sw_gamma2=sw_g2, sw_normal=sw_n, sw_packaged=sw_pk, sw_corr=sw_corr)
    np.savez_compressed(os.path.join(HERE, "upwind_penalty.npz"), **out)
    print("wrote upwind_penalty.npz")
    moving_mesh(gh, sw, out)


def moving_mesh(gh, sw, base):
    """tests/golden/upwind_penalty_moving.npz: dg_package_data of the same inputs with a
    normal_dot_mesh_velocity (the reference's twins subtract it from the characteristic
    speeds, GH with the factor 1 + gamma1 on the first one)."""
    rng = np.random.default_rng(777)
    npts = base["gh_u"].shape[1]

    def unpack_aa(v):
        t = np.zeros((4, 4))
        for a in range(4):
            for b in range(a, 4):
                t[a, b] = t[b, a] = v[sym4(a, b)]
        return t

    def unpack_iaa(v):
        t = np.zeros((3, 4, 4))
        for a in range(4):
            for b in range(a, 4):
                for i in range(3):
                    t[i, a, b] = t[i, b, a] = v[i + 3 * sym4(a, b)]
        return t
    gh_ndotv = rng.uniform(-0.8, 0.8, npts)
    sw_ndotv = rng.uniform(-1.5, 1.5, npts)
    gh_pk = np.zeros((npts, 134))
    sw_pk = np.zeros((npts, 16))
    for p in range(npts):
        u = base["gh_u"][0, p]
        r = gh.dg_package_data(unpack_aa(u[:10]), unpack_aa(u[10:20]), unpack_iaa(u[20:]),
                               base["gh_gamma1"][0, p], base["gh_gamma2"][0, p],
                               base["gh_lapse"][0, p], base["gh_shift"][0, p],
                               base["gh_nlo"][0, p], base["gh_nup"][0, p],
                               np.zeros(3), gh_ndotv[p])
        gh_pk[p] = np.concatenate([pack_aa(r[0]), pack_iaa(r[1]), pack_aa(r[2]), pack_aa(r[3]),
                                   pack_iaa(r[4]), pack_iaa(r[5]), pack_aa(r[6]), r[7]])
        # ScalarWave: the twin's moving-mesh branch does not run (its einsum("ijj->i", ...) has
        # three operands for one subscript list); all of its other outputs are speed * field
        # and are taken from it with the v^0 entry, the only one that branch computes, left out
        w = base["sw_u"][0, p]
        try:
            r = sw.dg_package_data(w[0], w[1], w[2:], base["sw_gamma2"][0, p],
                                   base["sw_normal"][0, p], np.zeros(3), sw_ndotv[p])
            sw_pk[p] = np.concatenate([[r[0]], r[1], [r[2]], [r[3]], r[4], r[5], [r[6]], r[7]])
        except ValueError:
            sw_pk[p] = np.nan
    np.savez_compressed(os.path.join(HERE, "upwind_penalty_moving.npz"), gh_ndotv=gh_ndotv,
                        gh_packaged=gh_pk, sw_ndotv=sw_ndotv, sw_packaged=sw_pk)
    print("wrote upwind_penalty_moving.npz")


if __name__ == "__main__":
    main()


def damped_harmonic():
    """tests/golden/damped_harmonic.npz from the reference's
    tests/Unit/Evolution/Systems/GeneralizedHarmonic/GaugeSourceFunctions/
    DampedHarmonic.py (exponents are hard-coded to 4 there).  The reference
    functions modify their metric argument in place (g_00 -= 1, g_ii += 1); the
    fixture stores the metric AFTER that shift, i.e. the metric the formulas see."""
    import sys
    sys.path.insert(0, "/root/reference/tests/Unit")
    from Evolution.Systems.GeneralizedHarmonic.GaugeSourceFunctions import DampedHarmonic as dh
    rng = np.random.default_rng(4242)
    npts = 48
    G = np.zeros((npts, 4, 4)); PI = np.zeros((npts, 4, 4)); PHI = np.zeros((npts, 3, 4, 4))
    X = np.zeros((npts, 3)); PRM = np.zeros((npts, 4)); Hs = np.zeros((npts, 4))
    dHs = np.zeros((npts, 4, 4))
    for p in range(npts):
        g = rng.uniform(-0.1, 0.1, (4, 4)); g = 0.5 * (g + g.T)
        pi = rng.uniform(-0.5, 0.5, (4, 4)); pi = 0.5 * (pi + pi.T)
        phi = rng.uniform(-0.5, 0.5, (3, 4, 4)); phi = 0.5 * (phi + phi.transpose(0, 2, 1))
        x = rng.uniform(-3, 3, 3)
        aL1, aL2, aS = rng.uniform(0.5, 2, 3)
        sigma = rng.uniform(5, 20)
        g_in = g.copy()
        H = dh.damped_harmonic_gauge_source_function(g_in, pi, phi, x, aL1, aL2, aS, sigma)
        g_in2 = g.copy()
        dH = dh.spacetime_deriv_damped_harmonic_gauge_source_function(g_in2, pi, phi, x, aL1, aL2,
                                                                      aS, sigma)
        assert np.array_equal(g_in, g_in2)
        G[p], PI[p], PHI[p], X[p] = g_in, pi, phi, x
        PRM[p] = (sigma, aL1, aL2, aS)
        Hs[p], dHs[p] = H, dH
    np.savez_compressed(os.path.join(HERE, "damped_harmonic.npz"), g=G, pi=PI, phi=PHI, x=X,
                        params=PRM, H=Hs, dH=dHs)
    print("wrote damped_harmonic.npz")


if __name__ == "__main__":
    damped_harmonic()
