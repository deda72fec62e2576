"""Non-conforming (2:1 h-refined) mortars on the CPU: the projection matrices of
Spectral/Projection.cpp (oracle restatement of the reference's closed forms vs
the library's independent computation), their defining properties as tested by
tests/Unit/NumericalAlgorithms/Spectral/Test_Projection.cpp, and the oracle's
mortar path (InternalMortarDataImpl.hpp:230-320, ApplyBoundaryCorrections.hpp:
797-1045) on an h-refined Brick."""
import os
from collections import Counter

import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain, lib

SIZES = (orc.MORTAR_LOWER_HALF, orc.MORTAR_UPPER_HALF)


@pytest.mark.parametrize("N", [2, 3, 5, 8, 12])
def test_projection_matrices(N):
    x, w = orc.lgl_points_and_weights(N)
    total = np.zeros((N, N))
    for size in SIZES:
        P = orc.projection_matrix_parent_to_child(N, N, size)
        R = orc.projection_matrix_child_to_parent(N, N, size)
        # the library computes both differently (barycentric formula; exact
        # quadrature of the L2 projection instead of the closed-form recurrence)
        np.testing.assert_allclose(lib.projection_matrix(N, False, size), P, atol=1e-13)
        np.testing.assert_allclose(lib.projection_matrix(N, True, size), R, atol=2e-13)
        # prolongation is exact for polynomials of the parent space
        # (Test_Projection.cpp:140-165)
        xc = 0.5 * (x + (1.0 if size == orc.MORTAR_UPPER_HALF else -1.0))
        for k in range(N):
            np.testing.assert_allclose(P @ x ** k, xc ** k, atol=1e-13)
        # restriction is the L2 projection of (f on the child's half, 0 on the other
        # half) onto the parent space: the error is orthogonal to every parent
        # polynomial over the whole parent interval (Test_Projection.cpp:230-260)
        rng = np.random.default_rng(N + size)
        f_child = rng.uniform(-1, 1, N)
        xg, wg = np.polynomial.legendre.leggauss(N + 1)
        sgn = 1.0 if size == orc.MORTAR_UPPER_HALF else -1.0
        x_here, x_other = 0.5 * (xg + sgn), 0.5 * (xg - sgn)   # parent coordinates
        f_here = orc.interpolation_matrix(N, xg) @ f_child     # child nodal -> Gauss points
        proj = R @ f_child
        p_here = orc.interpolation_matrix(N, x_here) @ proj
        p_other = orc.interpolation_matrix(N, x_other) @ proj
        for k in range(N):
            residual = np.sum(0.5 * wg * (f_here - p_here) * x_here ** k) + \
                np.sum(0.5 * wg * (0.0 - p_other) * x_other ** k)
            assert abs(residual) < 1e-12
        total += R @ P
    # restricting the prolongation of both halves gives the parent back
    # (Test_Projection.cpp:395-442)
    np.testing.assert_allclose(total, np.eye(N), atol=1e-12)
    np.testing.assert_array_equal(lib.projection_matrix(N, True, orc.MORTAR_FULL), np.eye(N))
    # lower and upper half are mirror images (Projection.cpp:246-253)
    np.testing.assert_allclose(orc.projection_matrix_child_to_parent(N, N, 1),
                               orc.projection_matrix_child_to_parent(N, N, 2)[::-1, ::-1])


def _poly(x):
    X, Y, Z = x[:, 0], x[:, 1], x[:, 2]
    psi = 1 + X + 2 * Y ** 2 - Z ** 3 + X * Y * Z
    pi = 0.5 - X ** 2 + Y * Z
    phi = [1 + Y * Z, 4 * Y + X * Z, -3 * Z ** 2 + X * Y]
    return np.stack([psi, pi] + phi, axis=1)


def test_refined_brick_tables():
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], 3, [(0, 0, 0), (1, 1, 0)])
    nb, mt = rb.neighbors(), rb.mortars()
    assert rb.n_elements == 6 + 16 and len(set(rb.element_ids())) == rb.n_elements
    coarse = Counter((m[0], m[1]) for m in mt)
    fine = Counter((m[2], m[3]) for m in mt)
    assert all(v == 4 for v in coarse.values()) and all(v == 1 for v in fine.values())
    assert len(coarse) + len(fine) == (nb == domain.HANGING).sum()
    x = rb.coords()
    for ec, dc, ef, df, sa, sb in mt:
        assert nb[ec, dc] == domain.HANGING and nb[ef, df] == domain.HANGING and df == dc ^ 1
        # the fine face is the stated quarter of the coarse face
        fc = x[ec][:, domain._face_point_indices(3, dc)]
        ff = x[ef][:, domain._face_point_indices(3, df)]
        fd = [d for d in range(3) if d != dc // 2]
        for dim, size in zip(fd, (sa, sb)):
            lo, hi = fc[dim].min(), fc[dim].max()
            mid = 0.5 * (lo + hi)
            want = (lo, mid) if size == domain.MORTAR_LOWER_HALF else (mid, hi)
            assert ff[dim].min() == pytest.approx(want[0]) and ff[dim].max() == pytest.approx(want[1])
    # conforming entries are symmetric
    for e in range(rb.n_elements):
        for d in range(6):
            if nb[e, d] >= 0:
                assert nb[nb[e, d], d ^ 1] == e


@pytest.mark.parametrize("N", [3, 5])
def test_oracle_mortars_consistency(N):
    """(1) Continuous polynomial data of the element's own degree: projection to
    the mortar is exact and every boundary correction vanishes.  (2) A jump that
    is a polynomial on the coarse face: the coarse element's correction, projected
    back from its four mortars, equals the conforming correction of an unrefined
    mesh with the same face data (the L2 projection reproduces polynomials of
    the parent space)."""
    refined = [(0, 0, 0)]
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N, refined, periodic=(False,) * 3)
    x, J, nb, mt = rb.coords(), rb.inverse_jacobian(), rb.neighbors(), rb.mortars()
    stat = np.full((rb.n_elements, 1, N ** 3), 0.3)
    u = _poly(x) if N >= 4 else _poly(x) * 0 + np.stack(
        [1 + x[:, 0] + x[:, 1] * x[:, 2], 0.5 - x[:, 1], 1 + x[:, 2], 4 * x[:, 1],
         x[:, 0] * x[:, 1]], axis=1)
    full = orc.dg_rhs(0, N, u, J, stat, nb, mortars=mt)
    vol = orc.dg_rhs(0, N, u, J, stat, nb, volume_only=True)
    assert np.max(np.abs(full - vol)) < 1e-12
    # (2) unrefined reference mesh; add a polynomial offset to every element that
    # lies inside the coarse cell (0,0,0) -> a polynomial jump across its faces
    base = domain.Brick([0, 0, 0], [1, 1, 1], [1, 1, 1], N, periodic=(False,) * 3)
    xb, Jb, nbb = base.coords(), base.inverse_jacobian(), base.neighbors()
    statb = np.full((base.n_elements, 1, N ** 3), 0.3)

    def data(xx, inside):
        v = np.stack([1 + xx[:, 0] + xx[:, 1] * xx[:, 2], 0.5 - xx[:, 1], 1 + xx[:, 2],
                      4 * xx[:, 1], xx[:, 0] * xx[:, 1]], axis=1)
        off = np.stack([0.3 + 0.2 * xx[:, 1], -0.1 + xx[:, 2] * 0.5, 0.2 * xx[:, 0],
                        0.1 + 0.0 * xx[:, 0], -0.3 * xx[:, 1]], axis=1)
        return v + off * inside[:, None, None]
    in_ref = np.array([c == (0, 0, 0) for c, ch in rb.elements], float)
    in_base = np.array([c == (0, 0, 0) for c in base.cells], float)
    r_ref = orc.dg_rhs(0, N, data(x, in_ref), J, stat, nb, mortars=mt)
    r_base = orc.dg_rhs(0, N, data(xb, in_base), Jb, statb, nbb)
    # compare on the coarse neighbours of the refined cell (they exist in both meshes)
    checked = 0
    for e, (c, ch) in enumerate(rb.elements):
        if ch is None and (nb[e] == domain.HANGING).any():
            eb = base.index_of[c]
            np.testing.assert_allclose(r_ref[e], r_base[eb], atol=1e-11)
            checked += 1
    assert checked == 3


def test_anisotropic_refinement_mortars():
    """Cells split in one or two dimensions only: mortars that are Full in one face
    dimension and a half in the other; continuous polynomial data still give no
    boundary correction, and the sizes describe the geometry."""
    N = 4
    split = {(0, 0, 0): (True, False, False), (1, 1, 1): (True, True, False),
             (1, 0, 0): (False, False, True)}
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N, split, periodic=(False,) * 3)
    nb, mt = rb.neighbors(), rb.mortars()
    assert rb.n_elements == 5 + 2 + 4 + 2 and len(set(rb.element_ids())) == rb.n_elements
    sizes = Counter((int(m[4]), int(m[5])) for m in mt)
    assert sizes[(1, 0)] and sizes[(0, 1)] and sizes[(1, 1)] and (0, 0) not in sizes
    x, J = rb.coords(), rb.inverse_jacobian()
    for ec, dc, ef, df, sa, sb in mt:
        fc = x[ec][:, domain._face_point_indices(N, dc)]
        ff = x[ef][:, domain._face_point_indices(N, df)]
        fd = [d for d in range(3) if d != dc // 2]
        for dim, size in zip(fd, (sa, sb)):
            lo, hi = fc[dim].min(), fc[dim].max()
            mid = 0.5 * (lo + hi)
            want = {0: (lo, hi), 1: (lo, mid), 2: (mid, hi)}[size]
            assert ff[dim].min() == pytest.approx(want[0]) and ff[dim].max() == pytest.approx(want[1])
    stat = np.full((rb.n_elements, 1, N ** 3), 0.3)
    u = _poly(x)
    full = orc.dg_rhs(0, N, u, J, stat, nb, mortars=mt)
    vol = orc.dg_rhs(0, N, u, J, stat, nb, volume_only=True)
    assert np.max(np.abs(full - vol)) < 1e-12
    with pytest.raises(NotImplementedError, match="smaller than both faces"):
        domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N,
                            {(0, 0, 0): (False, True, False), (1, 0, 0): (False, False, True)},
                            periodic=(False,) * 3).neighbors()


@pytest.mark.parametrize("system", ["sw", "gh"])
def test_oracle_mortars_between_non_aligned_blocks(system):
    """Every element of an h-refined periodic Brick (isotropic and anisotropic
    splits) gets its own rotated / reflected logical frame: neighbour directions,
    face permutations, oriented mortar rows and mortar sizes in the rotated coarse
    frame.  The right-hand side mapped back equals that of the aligned mesh."""
    from spectre_b200 import analytic
    from tests import rotation
    N = 4
    rng = np.random.default_rng(5)
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N,
                             {(0, 0, 0): (1, 1, 1), (1, 1, 0): (1, 0, 1), (0, 1, 1): (0, 0, 1)})
    x, nb, mt = rb.coords(), rb.neighbors(), rb.mortars()
    E = rb.n_elements
    J = rb.inverse_jacobian() + 0.05 * rng.uniform(-1, 1, (E, 9, N ** 3))
    if system == "gh":
        sid = 1
        u = analytic.gauge_wave(x, 0.1) + 1e-2 * rng.uniform(-1, 1, (E, 50, N ** 3))
        stat = rng.uniform(-1, 1, (E, 3, N ** 3))
        blocks = [slice(0, 10), slice(10, 20), slice(20, 50)]
    else:
        sid = 0
        u = analytic.plane_wave(x, 0.3) + 0.1 * rng.uniform(-1, 1, (E, 5, N ** 3))
        stat = rng.uniform(0, 1, (E, 1, N ** 3))
        blocks = [slice(0, 1), slice(1, 2), slice(2, 5)]
    all48 = rotation.signed_perms()
    frames = [all48[k] for k in rng.choice(48, E)]
    u_r, J_r, s_r, nbr_r, nd_r, perm_r, pm, mt_r = rotation.rotate_problem(
        N, u, J, stat, nb, frames, mortars=mt)
    assert ((mt_r[:, 3] >> 3) != 0).any() and ((mt_r[:, 3] & 7) != (mt_r[:, 1] ^ 1)).any()
    assert (mt_r[:, 4:] != np.asarray(mt)[:, 4:]).any()
    ref = orc.dg_rhs(sid, N, u, J, stat, nb, mortars=mt)
    got_r = orc.dg_rhs(sid, N, u_r, J_r, s_r, nbr_r, nbr_dir=nd_r, face_perm=perm_r,
                       mortars=mt_r)
    got = np.empty_like(got_r)
    for e in range(E):
        got[e] = got_r[e][:, pm[e]]
    err = max(np.max(np.abs(got[:, b] - ref[:, b])) / np.max(np.abs(ref[:, b])) for b in blocks)
    assert err < 1e-12, err


@pytest.mark.parametrize("n_parent,n_child", [(2, 3), (3, 5), (4, 4), (5, 8), (7, 12), (11, 12)])
def test_projection_matrices_between_different_meshes(n_parent, n_child):
    """p-refinement (the mortar mesh has more points than the element face):
    dgrhs_projection_matrix_meshes (barycentric interpolation; exact quadrature of the
    L2 projection) against the oracle's restatement of the closed forms of
    Projection.cpp:57-362, plus the properties Test_Projection.cpp checks."""
    xp, _ = orc.lgl_points_and_weights(n_parent)
    xc, _ = orc.lgl_points_and_weights(n_child)
    total = np.zeros((n_parent, n_parent))
    for size in (orc.MORTAR_FULL,) + SIZES:
        P = lib.projection_matrix_meshes(n_parent, n_child, False, size)
        R = lib.projection_matrix_meshes(n_parent, n_child, True, size)
        assert P.shape == (n_child, n_parent) and R.shape == (n_parent, n_child)
        np.testing.assert_allclose(P, orc.projection_matrix_parent_to_child(n_parent, n_child, size),
                                   atol=1e-13)
        np.testing.assert_allclose(R, orc.projection_matrix_child_to_parent(n_child, n_parent, size),
                                   atol=5e-13)
        t = {orc.MORTAR_FULL: xc, orc.MORTAR_UPPER_HALF: 0.5 * (xc + 1.0),
             orc.MORTAR_LOWER_HALF: 0.5 * (xc - 1.0)}[size]
        for k in range(n_parent):      # exact for the parent's polynomials
            np.testing.assert_allclose(P @ xp ** k, t ** k, atol=1e-13)
        if size == orc.MORTAR_FULL:    # projecting the interpolant gives the function back
            np.testing.assert_allclose(R @ P, np.eye(n_parent), atol=1e-12)
        else:
            total += R @ P
    np.testing.assert_allclose(total, np.eye(n_parent), atol=1e-12)
    with pytest.raises(lib.DgrhsError, match="n_parent <= n_child"):
        lib.projection_matrix_meshes(n_child + 1, n_child, False, 0)


def test_cpp_orientation_map_and_mortar_size_shims():
    """OrientationMap<3> -> (neighbour direction, face permutation) and dg::mortar_size of
    SpectreShims.hpp (what a caller uses to turn Element<3>::neighbors() into the tables
    of the C-ABI; host-only): on an h-refined Brick whose elements all carry random
    rotated / reflected frames, the relative OrientationMap of two neighbours is known
    from their frames; the shim must reproduce the direction / permutation tables found
    by brute-force point matching and the oriented mortar rows (tests/rotation.py)."""
    import subprocess
    from tests import rotation
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "orientation_codes")
    src = os.path.join(root, "tests", "helpers", "orientation_codes.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-o", exe, src, "-L",
                           os.path.join(root, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(root, "spectre_b200")])
    N = 3
    rng = np.random.default_rng(17)
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N,
                             {(0, 0, 0): (1, 1, 1), (1, 1, 0): (1, 0, 1), (0, 1, 1): (0, 0, 1)})
    nb, mt = rb.neighbors(), rb.mortars()
    E = rb.n_elements
    all48 = rotation.signed_perms()
    frames = [all48[k] for k in rng.choice(48, E)]
    z = np.zeros((E, 1, N ** 3))
    _, _, _, nbr_r, nd_r, perm_r, _, mt_r = rotation.rotate_problem(
        N, z, np.zeros((E, 9, N ** 3)), z, nb, frames, mortars=mt)

    def relative(e, o):
        """host (e) upper-a direction -> (dimension, sign) in the frame of o"""
        (pe, se), (po, so) = frames[e], frames[o]
        out = []
        for a in range(3):
            a2 = po.index(pe[a])
            out += [a2, se[a] * so[a2]]
        return out

    def segments(e):
        """(level, index) per rotated dimension"""
        c, ch = rb.elements[e]
        m = rb._mask(c)
        chh = ch or (0, 0, 0)
        lev = [rb.levels[d] + (1 if m[d] else 0) for d in range(3)]
        idx = [2 * c[d] + chh[d] if m[d] else c[d] for d in range(3)]
        p, s = frames[e]
        out = []
        for a in range(3):
            i = idx[p[a]] if s[a] > 0 else 2 ** lev[p[a]] - 1 - idx[p[a]]
            out += [lev[p[a]], i]
        return out
    lines, want = [], []
    for e in range(E):
        for d in range(6):
            o = nbr_r[e, d]
            if o >= 0:
                lines.append("F " + " ".join(map(str, relative(e, o) + [d])))
                want.append((int(nd_r[e, d]), int(perm_r[e, d])))
    for ec, dc, ef, dfp, sa, sb in mt_r.tolist():
        lines.append("F " + " ".join(map(str, relative(ec, ef) + [dc])))
        want.append((dfp & 7, dfp >> 3))
        lines.append("M " + " ".join(map(str, relative(ec, ef) + [dc // 2] + segments(ec)
                                         + segments(ef))))
        want.append((sa, sb))
        # seen from the fine element the mortar is its whole face
        lines.append("M " + " ".join(map(str, relative(ef, ec) + [(dfp & 7) // 2] + segments(ef)
                                         + segments(ec))))
        want.append((0, 0))
    # the known answers of Test_MortarHelpers.cpp:57-68 (2-d cases embedded in 3-d) and
    # :129-135 (non-aligned blocks: xi -> +eta, eta -> +zeta, zeta -> -xi)
    aligned = [0, 1, 1, 1, 2, 1]
    lines.append("M " + " ".join(map(str, aligned + [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])))
    want.append((0, 0))
    lines.append("M " + " ".join(map(str, aligned + [1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])))
    want.append((0, 0))
    lines.append("M " + " ".join(map(str, aligned + [1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0])))
    want.append((1, 0))
    lines.append("M 1 1 2 1 0 -1 0  0 0 3 2 7 5  6 61 3 0 4 5")
    want.append((2, 0))
    # OrientationMap known answers of Test_OrientationMap.cpp: :308-314 the inverse of
    # (-eta, -zeta, +xi) is (+zeta, -xi, -eta); :268-285 the all-flipped map takes the segments
    # (2,1) (3,1) (3,3) to (2,2) (3,6) (3,4) and is not aligned
    lines.append("I 1 -1 2 -1 0 1  0 0 0 0 0 0")
    want.append((2, 1, 0, -1, 1, -1, 0, 0, 0, 0, 0, 0, 0))
    lines.append("I 0 -1 1 -1 2 -1  2 1 3 1 3 3")
    want.append((0, -1, 1, -1, 2, -1, 2, 2, 3, 6, 3, 4, 0))
    lines.append("I 0 1 1 1 2 1  2 1 3 1 3 3")
    want.append((0, 1, 1, 1, 2, 1, 2, 1, 3, 1, 3, 3, 1))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = [tuple(map(int, ln.split())) for ln in out.stdout.strip().splitlines()]
    assert len(got) == len(want) > 100
    assert got == want


@pytest.mark.parametrize("periodic,rotated", [(True, True), (False, True), (False, False)])
def test_cpp_connectivity_builder(periodic, rotated):
    """DgConnectivity of SpectreShims.hpp: fed Element<3>-style neighbour lists (ids of the
    neighbours per direction + the OrientationMap to them) of an h-refined Brick whose
    elements all sit in random rotated frames, it must produce the neighbour table, the
    orientation tables and the oriented mortar rows of tests/rotation.py."""
    import subprocess
    from tests import rotation
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "orientation_codes")
    src = os.path.join(root, "tests", "helpers", "orientation_codes.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-o", exe, src, "-L",
                           os.path.join(root, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(root, "spectre_b200")])
    N = 2
    rng = np.random.default_rng(23 + periodic)
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N,
                             {(0, 0, 0): (1, 1, 1), (1, 1, 0): (1, 0, 1), (0, 1, 1): (0, 0, 1)},
                             periodic=(periodic,) * 3)
    nb, mt = rb.neighbors(), rb.mortars()
    E = rb.n_elements
    all48 = rotation.signed_perms()
    frames = [all48[k] if rotated else ((0, 1, 2), (1, 1, 1)) for k in rng.choice(48, E)]
    z = np.zeros((E, 1, N ** 3))
    _, _, _, nbr_r, nd_r, perm_r, _, mt_r = rotation.rotate_problem(
        N, z, np.zeros((E, 9, N ** 3)), z, nb, frames, mortars=mt)

    def relative(e, o):
        (pe, se), (po, so) = frames[e], frames[o]
        out = []
        for a in range(3):
            a2 = po.index(pe[a])
            out += [a2, se[a] * so[a2]]
        return out

    def segments(e):
        c, ch = rb.elements[e]
        m = rb._mask(c)
        chh = ch or (0, 0, 0)
        lev = [rb.levels[d] + (1 if m[d] else 0) for d in range(3)]
        idx = [2 * c[d] + chh[d] if m[d] else c[d] for d in range(3)]
        p, s = frames[e]
        return [v for a in range(3)
                for v in (lev[p[a]], idx[p[a]] if s[a] > 0 else 2 ** lev[p[a]] - 1 - idx[p[a]])]
    lines = [f"C {E}"] + [" ".join(map(str, segments(e))) for e in range(E)]
    fine_of = {}      # (coarse, direction) -> fine elements
    coarse_of = {}    # (fine, direction) -> coarse element
    for ec, dc, ef, dfp, _, _ in mt_r.tolist():
        fine_of.setdefault((ec, dc), []).append(ef)
        coarse_of[(ef, dfp & 7)] = ec
    for e in range(E):
        for d in range(6):
            v = int(nbr_r[e, d])
            if v >= 0:
                others = [v]
            elif v == domain.HANGING:
                others = fine_of.get((e, d)) or [coarse_of[(e, d)]]
            else:
                lines.append(f"E {e} {d} {v}")
                continue
            # (every element of this mesh has its own frame, so the finer neighbours of a
            # face are handed over one by one, each with its own OrientationMap)
            if rotated:
                for o in others:
                    lines.append(" ".join(map(str, ["N", e, d, 1, o] + relative(e, o))))
            else:   # one block: the whole list with the block's (aligned) OrientationMap
                lines.append(" ".join(map(str, ["N", e, d, len(others)] + others
                                          + relative(e, others[0]))))
    lines.append("end")
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    tables = {ln.split()[0]: np.array(ln.split()[1:], dtype=np.int64)
              for ln in out.stdout.strip().splitlines()}
    np.testing.assert_array_equal(tables["neighbors"].reshape(E, 6), nbr_r)
    conforming = nbr_r >= 0
    np.testing.assert_array_equal(tables["directions"].reshape(E, 6)[conforming], nd_r[conforming])
    np.testing.assert_array_equal(tables["permutations"].reshape(E, 6)[conforming],
                                  perm_r[conforming])
    got_rows = sorted(map(tuple, tables["mortars"].reshape(-1, 6).tolist()))
    assert got_rows == sorted(map(tuple, mt_r.tolist())) and len(got_rows) > 0
    assert tables["aligned"][0] == (0 if rotated else 1)
    assert periodic or (nbr_r == -1).any()
    if not rotated:
        np.testing.assert_array_equal(nbr_r, nb)
        assert sorted(map(tuple, mt_r.tolist())) == sorted(map(tuple, np.asarray(mt).tolist()))


def test_oracle_p_mortar_path_reduces_to_conforming():
    """dg_rhs_p_refined (numpy: packaged data per face, projection to the mortar mesh,
    boundary terms, projection back, lift) with two classes of EQUAL N is the conforming
    right-hand side of the C oracle: pins the face normal, packaging, link and lift
    conventions of the p-mortar oracle path (the projections are pinned to
    Test_Projection.cpp's closed forms in test_oracle_pins.py)."""
    from spectre_b200 import analytic, domain
    for system in (0, 1):
        N = 4
        brick = domain.Brick([0, 0, 0], [1.0, 0.5, 0.5], [1, 0, 0], N)
        x, J, nb = brick.coords(), brick.inverse_jacobian(), brick.neighbors()
        rng = np.random.default_rng(system)
        if system == 0:
            u = analytic.plane_wave(x * 2 * np.pi, 0.1) + 1e-2 * rng.uniform(-1, 1, (2, 5, N ** 3))
            stat = rng.uniform(0.5, 1.5, (2, 1, N ** 3))
        else:
            u = analytic.gauge_wave(x, 0.05) + 1e-3 * rng.uniform(-1, 1, (2, 50, N ** 3))
            stat = np.zeros((2, 3, N ** 3))
            stat[:, 0], stat[:, 1], stat[:, 2] = 1.0, -1.0, rng.uniform(0.5, 1.5, (2, N ** 3))
        ref = orc.dg_rhs(system, N, u, J, stat, nb)
        classes = [{"N": N, "u": u[k:k + 1], "invjac": J[k:k + 1], "static": stat[k:k + 1],
                    "nbr": np.array([[orc.P_MORTAR, orc.P_MORTAR, 0, 0, 0, 0]], dtype=np.int32)}
                   for k in range(2)]
        links = [(0, 0, 1, 1, 0, 0, 0), (0, 0, 0, 1, 0, 1, 0)]
        got = orc.dg_rhs_p_refined(system, classes, links)
        for k in range(2):
            assert np.max(np.abs(got[k] - ref[k:k + 1])) < 1e-14 * np.max(np.abs(ref))


def test_oracle_p_mortar_projection_properties():
    """A p-mortar between N = 4 and N = 6: data of the coarse side's polynomial degree are
    reproduced exactly by project_to_mortar followed by project_from_mortar, and the face
    orientation maps are permutations whose inverse is again one of the eight codes."""
    P = orc.projection_matrix_parent_to_child(4, 6, orc.MORTAR_FULL)
    R = orc.projection_matrix_child_to_parent(6, 4, orc.MORTAR_FULL)
    assert np.max(np.abs(R @ P - np.eye(4))) < 1e-13
    for perm in range(8):
        m = orc.orient_face_map(5, perm)
        assert sorted(m.tolist()) == list(range(25))
        assert any(np.array_equal(orc.orient_face_map(5, q)[m], np.arange(25)) for q in range(8))
