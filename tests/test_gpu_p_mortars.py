"""p-refinement: elements with different numbers of grid
points live in one context per N; the faces between them are p-mortars (mortar mesh = the
larger extents, MortarHelpers.cpp:22-49).  GPU (pmortar_kernel through the C-ABI) vs the
oracle's dg_rhs_p_refined, which projects the reference's packaged data with the projection
matrices pinned in tests/test_oracle_pins.py."""
import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, lib

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _two_class_problem(system, NA, NB, perm, seed):
    """two elements side by side in x, periodic in all directions: element 0 has NA points per
    dimension, element 1 has NB; the y / z neighbours are the elements themselves"""
    classes = []
    rng = np.random.default_rng(seed)
    for k, N in enumerate((NA, NB)):
        brick = domain.Brick([0, 0, 0], [1.0, 0.5, 0.5], [1, 0, 0], N)
        x, J = brick.coords()[k:k + 1], brick.inverse_jacobian()[k:k + 1]
        if system == lib.SYSTEM_SCALAR_WAVE:
            u = analytic.plane_wave(x * 2 * np.pi, 0.1) + 1e-2 * rng.uniform(-1, 1, (1, 5, N ** 3))
            stat = rng.uniform(0.5, 1.5, (1, 1, N ** 3))
        else:
            u = analytic.gauge_wave(x, 0.05) + 1e-3 * rng.uniform(-1, 1, (1, 50, N ** 3))
            stat = np.zeros((1, 3, N ** 3))
            stat[:, 0], stat[:, 1], stat[:, 2] = 1.0, -1.0, rng.uniform(0.5, 1.5, (1, N ** 3))
        nbr = np.array([[orc.P_MORTAR, orc.P_MORTAR, 0, 0, 0, 0]], dtype=np.int32)
        classes.append({"N": N, "u": u, "invjac": J, "static": stat, "nbr": nbr, "x": x})
    # upper x face of element 0 meets the lower x face of element 1 and vice versa
    links = [(0, 0, 1, 1, 0, 0, perm), (0, 0, 0, 1, 0, 1, 0)]
    return classes, links


def _inverse_perm(perm):
    for q in range(8):
        if np.array_equal(orc.orient_face_map(5, q)[orc.orient_face_map(5, perm)], np.arange(25)):
            return q
    raise AssertionError


def _gpu_contexts(system, classes, links):
    ctxs = []
    for cl in classes:
        ctx = lib.Context(system, cl["N"], 1, 2)
        ctx.set_geometry(cl["invjac"], cl["x"], cl["nbr"])
        ctx.set_static_fields(cl["static"])
        ctx.set_state(cl["u"])
        ctxs.append(ctx)
    tables = [[], []]
    sends = [[], []]
    for (ca, ea, da, cb, eb, db, perm) in links:
        tables[ca].append([ea, da, classes[cb]["N"], db | (perm << 3)])
        tables[cb].append([eb, db, classes[ca]["N"], da | (_inverse_perm(perm) << 3)])
        sends[ca].append([ea, da])
        sends[cb].append([eb, db])
    for k, ctx in enumerate(ctxs):
        ctx.set_p_mortars(tables[k])
        ctx.set_halo_map(np.array(sends[k], dtype=np.int32))
        ctx.set_interior_count(0)
    return ctxs


def _gpu_rhs(ctxs, time=0.0):
    for ctx in ctxs:
        ctx.pack_halo()
    # link i is slot i / face i on both sides
    n = ctxs[0]._n_links
    ctxs[1].p_mortar_transfer_from(ctxs[0], np.arange(n), np.arange(n))
    ctxs[0].p_mortar_transfer_from(ctxs[1], np.arange(n), np.arange(n))
    for ctx in ctxs:
        ctx.compute_time_derivative_range(time, 0, 1)
    return [ctx.get_time_derivative() for ctx in ctxs]


@pytest.mark.parametrize("system", [lib.SYSTEM_SCALAR_WAVE, lib.SYSTEM_GH])
@pytest.mark.parametrize("NA,NB,perm", [(5, 7, 0), (7, 5, 0), (4, 12, 0), (6, 5, 3), (3, 4, 5),
                                        (9, 8, 6), (11, 12, 7)])
def test_p_mortar_rhs(system, NA, NB, perm):
    classes, links = _two_class_problem(system, NA, NB, perm, 100 * NA + NB)
    ctxs = _gpu_contexts(system, classes, links)
    for ctx in ctxs:
        ctx._n_links = len(links)
    got = _gpu_rhs(ctxs)
    want = orc.dg_rhs_p_refined(system, classes, links)
    volume = [orc.dg_rhs(system, cl["N"], cl["u"], cl["invjac"], cl["static"], cl["nbr"] * 0 - 1)
              for cl in classes]
    for k in range(2):
        scale = np.max(np.abs(want[k]))
        assert np.max(np.abs(got[k] - want[k])) < TOL * scale
        # the p-mortar terms are a visible part of the right-hand side
        assert np.max(np.abs(want[k] - volume[k])) > 1e-6 * scale
    for ctx in ctxs:
        ctx.close()


def test_p_mortar_evolution_and_misuse():
    """Two contexts stepped together (AB3 incl. self-start): pack, transfer, RHS, update per
    substep, vs the oracle's Evolution on the concatenated state."""
    system, NA, NB, dt, steps = lib.SYSTEM_SCALAR_WAVE, 5, 7, 1e-3, 3
    classes, links = _two_class_problem(system, NA, NB, 0, 7)
    ctxs = _gpu_contexts(system, classes, links)
    for ctx in ctxs:
        ctx._n_links = len(links)
        ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
    done = 0
    while done < steps:
        times = [ctx.begin_substep() for ctx in ctxs]
        assert times[0] == times[1]
        _gpu_rhs(ctxs, times[0])
        done += [ctx.end_substep() for ctx in ctxs][0]
    sizes = [cl["u"].size for cl in classes]

    def rhs(v, t):
        for cl, part in zip(classes, np.split(v, [sizes[0]])):
            cl["u"] = part.reshape(cl["u"].shape)
        return np.concatenate([r.ravel() for r in orc.dg_rhs_p_refined(system, classes, links)])

    u0 = np.concatenate([cl["u"].ravel() for cl in classes])
    ev = orc.Evolution(rhs, u0, 0.0, dt, "AB3")
    for _ in range(steps):
        ev.step()
    got = np.concatenate([ctx.get_state().ravel() for ctx in ctxs])
    assert np.max(np.abs(got - ev.u)) < TOL * np.max(np.abs(ev.u))
    # misuse
    with pytest.raises(lib.DgrhsError, match="marked DGRHS_NEIGHBOR_P_MORTAR but the table"):
        ctxs[0].set_p_mortars([[0, 1, NB, 0]])
    with pytest.raises(lib.DgrhsError, match="equal N"):
        ctxs[0].set_p_mortars([[0, 1, NA, 0], [0, 0, NB, 1]])
    for ctx in ctxs:
        ctx.close()


def test_p_refined_binary_domain_matches_oracle():
    """Block groups of the BinaryCompactObject domain with different N (the isotropic part of
    Inspiral.yaml:102-108's InitialGridPoints): p-mortars between non-aligned curved blocks
    (wedge <-> cube wedge <-> frustum <-> half wedge), DirichletAnalytic on the three spheres,
    AnalyticChristoffel gauge.  RHS and two self-started AB3 steps vs the oracle."""
    from spectre_b200 import analytic, bco, p_refinement
    dom = bco.BinaryCompactObject(8.0, -8.0, 0.8, 4.0, 0.8, 4.0, 60.0, 300.0, 0, 4,
                                  opening_angle_degrees=120.0)
    points = {"ObjectAShell": 6, "ObjectACube": 4, "ObjectBShell": 6, "ObjectBCube": 5,
              "Envelope": 6, "OuterShell": 5}
    pts = [points[name] for name in dom.block_names]
    centers = ((8.0, 0.0, 0.0), (-8.0, 0.0, 0.0))
    data = lambda x, t: analytic.superposed_kerr_schild(x, (0.5, 0.5), centers)   # noqa: E731
    dt = 1e-3
    ev = p_refinement.PRefinedEvolution(lib.SYSTEM_GH, dom, pts, data, (1.0, -1.0, 1.0), dt=dt)
    assert ev.Ns == [4, 5, 6] and sum(len(i) for i in ev.ids) == 44
    rng = np.random.default_rng(5)
    classes = []
    for k, N in enumerate(ev.Ns):
        u = ev.u0[k] + 1e-3 * rng.uniform(-1, 1, ev.u0[k].shape)
        ev.ctxs[k].set_state(u)
        ev.ctxs[k].set_stepper(lib.STEPPER_ADAMS_BASHFORTH, 3, 0.0, dt)
        H = np.zeros((len(u), 4, N ** 3))
        dH = np.zeros((len(u), 16, N ** 3))
        for e in range(len(u)):
            H[e], dH[e] = orc.analytic_christoffel_gauge(N, ev.u0[k][e], ev.J[k][e])
        t = ev.tables[k]
        ext = ev.boundary_ghost_data(k, ev.x[k], ev.J[k], ev.stat[k], ev.u0[k])[:, :50]
        classes.append({"N": N, "u": u, "invjac": ev.J[k],
                        "static": np.concatenate([ev.stat[k], H, dH], axis=1),
                        "nbr": t["nbr"], "nbr_dir": t["nbr_dir"], "face_perm": t["face_perm"],
                        "ext_u": ext})
    # links from the tables: every interface once, seen from the class that lists it first
    links, seen = [], set()
    cls_of = {N: k for k, N in enumerate(ev.Ns)}
    local = {int(g): (k, le) for k, ids in enumerate(ev.ids) for le, g in enumerate(ids)}
    for a, t in enumerate(ev.tables):
        for (le, d, nb_points, code, g2) in t["pm"]:
            b, le2 = local[g2]
            key = tuple(sorted([(a, le, d), (b, le2, code & 7)]))
            if key in seen:
                continue
            seen.add(key)
            links.append((a, le, d, b, le2, code & 7, code >> 3))
    assert len(links) == sum(len(t["pm"]) for t in ev.tables) // 2 > 20
    ev.compute_time_derivative(0.0)
    got = [ctx.get_time_derivative() for ctx in ev.ctxs]
    want = orc.dg_rhs_p_refined(1, classes, links, gauge_params=orc.GAUGE_GIVEN)
    scale = max(np.max(np.abs(w)) for w in want)
    for g, w in zip(got, want):
        assert np.max(np.abs(g - w)) < TOL * scale
    ev.take_steps(2)
    sizes = np.cumsum([c["u"].size for c in classes])[:-1]

    def rhs(v, t):
        for c, part in zip(classes, np.split(v, sizes)):
            c["u"] = part.reshape(c["u"].shape)
        return np.concatenate([r.ravel() for r in orc.dg_rhs_p_refined(
            1, classes, links, gauge_params=orc.GAUGE_GIVEN)])
    o = orc.Evolution(rhs, np.concatenate([c["u"].ravel() for c in classes]), 0.0, dt, "AB3")
    o.step()
    o.step()
    state = np.concatenate([ctx.get_state().ravel() for ctx in ev.ctxs])
    assert np.max(np.abs(state - o.u)) < TOL * np.max(np.abs(o.u))
    ev.close()
