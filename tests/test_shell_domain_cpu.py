"""Spherical-shell domain (Sphere creator with excision): the Wedge<3> map, the
element geometry, the multi-block connectivity with non-aligned neighbours and
the oracle's handling of orientations -- all on the CPU.

Reference: Domain/CoordinateMaps/Wedge.cpp, Domain/DomainHelpers.cpp:553-578
(orientations_for_sphere_wrappings), Domain/Creators/Sphere.cpp,
Domain/Structure/OrientationMapHelpers.cpp:25-120; pins from
tests/Unit/Domain/CoordinateMaps/Test_Wedge3D.cpp:175-283."""
import numpy as np
import pytest

from oracle import oracle as orc
from spectre_b200 import analytic, domain

# all_wedge_directions(), tests/Unit/Helpers/Domain/CoordinateMaps/TestMapHelpers.hpp:871-892
TEST_WEDGE_DIRECTIONS = [
    ((0, 1), (1, 1), (2, 1)), ((0, 1), (1, -1), (2, -1)),
    ((1, 1), (2, 1), (0, 1)), ((1, 1), (2, -1), (0, -1)),
    ((2, 1), (0, 1), (1, 1)), ((2, -1), (0, -1), (1, 1)),
]


@pytest.mark.parametrize("equiangular", [True, False])
@pytest.mark.parametrize("distribution", ["Linear", "Logarithmic", "Inverse"])
def test_wedge_alignment_known_answers(equiangular, distribution):
    """test_wedge3d_alignment (Test_Wedge3D.cpp:175-283): where the logical axes
    and the lowest corner of the six wedges land (inner radius sqrt 3, outer
    radius 2 sqrt 3)."""
    r3 = np.sqrt(3.0)

    def m(w, p):
        return domain.wedge_map(*p, r3, 2 * r3, TEST_WEDGE_DIRECTIONS[w], equiangular,
                                distribution)[0]
    lowest, ax, ae, az = (-1, -1, -1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)
    # wedge: (component, value) reached along xi, eta, zeta; lowest physical corner
    expected = {
        0: ((0, 1.0), (1, 1.0), (2, 2.0), (-1, -1, 1)),     # upper zeta: +X, +Y, +Z
        2: ((2, 1.0), (0, 1.0), (1, 2.0), (-1, 1, -1)),     # upper eta:  +Z, +X, +Y
        4: ((1, 1.0), (2, 1.0), (0, 2.0), (1, -1, -1)),     # upper xi:   +Y, +Z, +X
        1: ((0, 1.0), (1, -1.0), (2, -2.0), (-1, 1, -1)),   # lower zeta: +X, -Y, -Z
        3: ((2, -1.0), (0, 1.0), (1, -2.0), (-1, -1, 1)),   # lower eta:  -Z, +X, -Y
        5: ((1, -1.0), (2, 1.0), (0, -2.0), (-1, 1, -1)),   # lower xi:   -Y, +Z, -X
    }
    for w, (a, b, c, corner) in expected.items():
        assert m(w, ax)[a[0]] == pytest.approx(a[1], abs=1e-13)
        assert m(w, ae)[b[0]] == pytest.approx(b[1], abs=1e-13)
        assert m(w, az)[c[0]] == pytest.approx(c[1], abs=1e-13)
        np.testing.assert_allclose(m(w, lowest), corner, atol=1e-13)


@pytest.mark.parametrize("distribution", ["Linear", "Logarithmic", "Inverse"])
def test_wedge_jacobian_and_radii(distribution):
    """The analytic Jacobian against centred differences (as test_jacobian /
    test_inv_jacobian in TestMapHelpers.hpp do), and |x| = r(zeta) on spherical
    wedges (test_wedge3d_random_radii, Test_Wedge3D.cpp:390-430)."""
    rng = np.random.default_rng(5)
    r_in, r_out = 1.7, 5.3
    for w in range(6):
        p = rng.uniform(-1, 1, 3)
        x, jac = domain.wedge_map(*p, r_in, r_out, w, True, distribution)
        h = 1e-6
        for j in range(3):
            dp = np.zeros(3)
            dp[j] = h
            xp = domain.wedge_map(*(p + dp), r_in, r_out, w, True, distribution)[0]
            xm = domain.wedge_map(*(p - dp), r_in, r_out, w, True, distribution)[0]
            np.testing.assert_allclose(jac[:, j], (xp - xm) / (2 * h), rtol=1e-8, atol=1e-8)
        for zeta, r in ((-1.0, r_in), (1.0, r_out)):
            x = domain.wedge_map(p[0], p[1], zeta, r_in, r_out, w, True, distribution)[0]
            assert np.linalg.norm(x) == pytest.approx(r, rel=1e-14)
        assert np.linalg.det(jac) > 0  # handedness is preserved by the six rotations


def test_shell_geometry_and_connectivity():
    """KerrSchild.yaml:80-98 (6 wedges, r in [1.9, 2.3]) refined once with two
    layers: every angular face has exactly one neighbour, the radial ones are
    the excision / outer boundary or the next layer, and the orientation table is
    made of inverse pairs."""
    N = 4
    sh = domain.SphericalShell(1.9, 2.5, (1, 1), N, radial_partitioning=(2.2,))
    assert sh.n_blocks == 12 and sh.n_elements == 12 * 8
    x, J = sh.coords(), sh.inverse_jacobian()
    r = np.sqrt((x ** 2).sum(axis=1))
    assert r.min() == pytest.approx(1.9, rel=1e-14) and r.max() == pytest.approx(2.5, rel=1e-14)
    nbr, nd, perm = sh.neighbors(), *sh.neighbor_orientations()
    assert (nbr[:, :4] >= 0).all()
    assert (nbr == -1).sum() == 2 * 6 * 4       # inner + outer sphere, 4 elements per wedge
    assert (perm != 0).any() and (nd != (np.arange(6) ^ 1)[None, :]).any()
    q = np.arange(N * N)
    qa, qb = q % N, q // N

    def point_map(code):
        na, nb_ = np.where(code & 1, qb, qa), np.where(code & 1, qa, qb)
        if code & 2:
            na = N - 1 - na
        if code & 4:
            nb_ = N - 1 - nb_
        return na + N * nb_
    for e in range(sh.n_elements):
        for d in range(6):
            v = nbr[e, d]
            if v < 0:
                assert d >= 4
                continue
            assert nbr[v, nd[e, d]] == e and nd[v, nd[e, d]] == d
            there = point_map(perm[e, d])
            back = point_map(perm[v, nd[e, d]])
            assert np.array_equal(back[there], q)
            # the coordinates of matched face points coincide
            mine = x[e][:, domain._face_point_indices(N, d)]
            theirs = x[v][:, domain._face_point_indices(N, nd[e, d])][:, there]
            np.testing.assert_allclose(mine, theirs, atol=1e-13)
    # element ids: block index in the low byte, refinement levels per dimension
    ids = sh.element_ids()
    assert len(set(ids)) == sh.n_elements and (ids[8] & 0xFF) == 1
    # the numerical derivative of the coordinates with the inverse Jacobian is the
    # identity up to the truncation error of the (non-polynomial) map
    errs = []
    for Np in (4, 8):
        shp = domain.SphericalShell(1.9, 2.5, (1, 1), Np, radial_partitioning=(2.2,))
        xe, Je = shp.coords([3, 40]), shp.inverse_jacobian([3, 40])
        err = 0.0
        for k in range(2):
            du = orc.partial_derivatives(Np, xe[k], Je[k])     # [3 comps][3 derivs][n]
            du = np.asarray(du).reshape(3, 3, -1)
            for c in range(3):
                for i in range(3):
                    err = max(err, np.max(np.abs(du[c, i] - (1.0 if c == i else 0.0))))
        errs.append(err)
    # spectral convergence: 1.2e-2 at N = 4, 8.2e-6 at N = 8
    assert errs[0] < 5e-2 and errs[1] < 5e-5 and errs[1] < errs[0] * 2e-3


def _rotated_problem():
    from tests.test_gpu_orientation import _rotate_problem, _signed_perms
    N = 3
    rng = np.random.default_rng(12)
    brick = domain.Brick([0, 0, 0], [1.0] * 3, [1, 1, 1], N)
    x, nbr = brick.coords(), brick.neighbors()
    J = brick.inverse_jacobian() + 0.1 * rng.uniform(-1, 1, (brick.n_elements, 9, N ** 3))
    u = analytic.gauge_wave(x, 0.1) + 1e-2 * rng.uniform(-1, 1, (brick.n_elements, 50, N ** 3))
    stat = rng.uniform(-1, 1, (brick.n_elements, 3, N ** 3))
    all48 = _signed_perms()
    frames = [all48[k] for k in rng.choice(48, brick.n_elements, replace=False)]
    return N, u, J, stat, nbr, _rotate_problem(N, u, J, stat, nbr, frames)


def test_oracle_orientation_matches_aligned_mesh():
    """orient_variables_on_slice in the oracle: the right-hand side on a mesh
    whose elements carry random ones of the 48 cube orientations equals the
    aligned one after mapping back (the data are inertial tensor components)."""
    N, u, J, stat, nbr, (u_r, J_r, s_r, nbr_r, nd_r, perm_r, pm) = _rotated_problem()
    ref = orc.dg_rhs(1, N, u, J, stat, nbr)
    got_r = orc.dg_rhs(1, N, u_r, J_r, s_r, nbr_r, nbr_dir=nd_r, face_perm=perm_r)
    got = np.empty_like(got_r)
    for e in range(u.shape[0]):
        got[e] = got_r[e][:, pm[e]]
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-13


def test_partition_of_shell_keeps_orientations():
    """Partition of a multi-block domain: ghost faces carry the neighbour's
    direction and permutation, send and receive lists agree pairwise."""
    sh = domain.SphericalShell(1.9, 2.3, (1, 0), 3)
    nbr, (nd, perm) = sh.neighbors(), sh.neighbor_orientations()
    world = 3
    parts = [domain.Partition(nbr, world, r, boundary_slots=True, neighbor_direction=nd,
                              face_permutation=perm) for r in range(world)]
    assert sum(p.n_local for p in parts) == sh.n_elements
    for r, p in enumerate(parts):
        assert p.oriented
        for peer in range(world):
            assert p.recv_counts[peer] == parts[peer].send_counts[r]
        for le, g in enumerate(p.global_ids):
            for d in range(6):
                if nbr[g, d] >= 0:
                    assert p.local_neighbor_direction[le, d] == nd[g, d]
                    assert p.local_face_permutation[le, d] == perm[g, d]
                else:
                    assert p.local_neighbors[le, d] <= -2   # DirichletAnalytic slot
                    assert p.local_neighbor_direction[le, d] == d ^ 1


def test_radial_element_order_cuts_the_shell_at_constant_radius():
    """order="radial": a contiguous partition of the element list cuts the shell
    into spherical layers; the halo of a rank is two spheres of element faces."""
    sh = domain.SphericalShell(1.9, 30.0, (1, 3), 3, order="radial")
    nbr, (nd, perm) = sh.neighbors(), sh.neighbor_orientations()
    assert sorted(sh.cells) == sorted(domain.SphericalShell(1.9, 30.0, (1, 3), 3).cells)
    world = 4
    for r in range(world):
        p = domain.Partition(nbr, world, r, boundary_slots=True, neighbor_direction=nd,
                             face_permutation=perm)
        assert p.n_local == 6 * 4 * 2
        assert p.n_recv == (1 if r in (0, world - 1) else 2) * 6 * 4
        radial = {sh.cells[g][1][2] for g in p.global_ids}
        assert radial == {2 * r, 2 * r + 1}


def test_shell_layers_of_different_refinement_have_mortars():
    """Per-layer refinement: the inner layer one angular level finer than the
    outer one -> every element face on the interface sphere is a 2:1 mortar."""
    sh = domain.SphericalShell(1.9, 2.9, [(2, 0), (1, 0)], 3, radial_partitioning=(2.3,))
    nb, mt = sh.neighbors(), sh.mortars()
    assert sh.n_elements == 6 * 16 + 6 * 4
    assert (nb == domain.HANGING).sum() == 6 * 16 + 6 * 4 and len(mt) == 6 * 16
    assert (nb == -1).sum() == 6 * 16 + 6 * 4      # excision sphere + outer sphere
    x = sh.coords()
    P = [np.eye(3), orc.projection_matrix_parent_to_child(3, 3, 1),
         orc.projection_matrix_parent_to_child(3, 3, 2)]
    for ec, dc, ef, df, sa, sb in mt:
        assert dc == 4 and df == 5       # the coarse (outer) layer looks inwards
        # interpolating the coarse face's coordinates to the mortar gives the fine
        # face's points up to the interpolation error of the (non-polynomial) map
        fc = x[ec][:, domain._face_point_indices(3, dc)].reshape(3, 3, 3)   # [xyz, b, a]
        ff = x[ef][:, domain._face_point_indices(3, df)].reshape(3, 3, 3)
        interp = np.einsum("Bb,Aa,xba->xBA", P[sb], P[sa], fc)
        assert np.max(np.abs(interp - ff)) < 2e-2
        np.testing.assert_allclose(np.linalg.norm(ff, axis=0), 2.3, rtol=1e-13)


@pytest.mark.parametrize("fine_wedge_refinement,n_split", [((1, 1), 4), ((1, 0), 2)])
def test_shell_wedges_of_different_refinement_have_oriented_mortars(fine_wedge_refinement,
                                                                    n_split):
    """Per-block InitialRefinement: one wedge finer than its four neighbours -> the
    wedge-to-wedge interfaces are 2:1 mortars between blocks that are NOT aligned
    (four quarter mortars, or two half mortars when only the angular level
    differs).  The mortar rows carry the fine face's permutation; the coarse face's
    coordinates, interpolated to the mortar in the coarse frame, land on the fine
    face's points after that permutation."""
    N = 5
    perms_seen, not_opposite = set(), False
    for fine_wedge in range(6):
        ref = [(0, 0)] * 6
        ref[fine_wedge] = fine_wedge_refinement
        sh = domain.SphericalShell(1.9, 2.9, [ref], N)
        nb, mt = sh.neighbors(), sh.mortars()
        n_fine = 4 * 2 ** fine_wedge_refinement[1]
        assert sh.n_elements == 5 + n_fine
        assert len(mt) == 4 * n_split and (nb == domain.HANGING).sum() == 4 + 4 * n_split
        not_opposite |= bool(((mt[:, 3] & 7) != (mt[:, 1] ^ 1)).any())
        perms_seen |= set((mt[:, 3] >> 3).tolist())
        x = sh.coords()
        P = [np.eye(N), orc.projection_matrix_parent_to_child(N, N, 1),
             orc.projection_matrix_parent_to_child(N, N, 2)]
        q = np.arange(N * N)
        a, b = q % N, q // N
        for ec, dc, ef, dfp, sa, sb in mt:
            df, perm = dfp & 7, dfp >> 3
            assert sh.cells[ec][0] != fine_wedge and sh.cells[ef][0] == fine_wedge
            fc = x[ec][:, domain._face_point_indices(N, dc)].reshape(3, N, N)   # [xyz, b, a]
            interp = np.einsum("Bb,Aa,xba->xBA", P[sb], P[sa], fc).reshape(3, N * N)
            fa, fb = (b, a) if perm & 1 else (a, b)
            if perm & 2:
                fa = N - 1 - fa
            if perm & 4:
                fb = N - 1 - fb
            ff = x[ef][:, domain._face_point_indices(N, df)][:, fa + N * fb]
            assert np.max(np.abs(interp - ff)) < 2e-3
            # a wrong permutation would be off by the size of the face
            wrong = x[ef][:, domain._face_point_indices(N, df)][:, (N - 1 - fa) + N * fb]
            assert np.max(np.abs(interp - wrong)) > 0.1
    # rotated and reflected interfaces occur (the +z wedge alone is aligned with its neighbours)
    assert len(perms_seen) >= 2 and not_opposite, perms_seen


def test_oracle_static_black_hole_on_wedge_refined_shell():
    """Exact Kerr-Schild data on a shell with one wedge refined: with the oriented
    mortar rows the right-hand side stays at the truncation level of the coarse
    wedges; with the permutation bits stripped the mortar data are mismatched and
    it is tens of times larger."""
    from spectre_b200 import evolution
    N = 5

    def max_rhs(refinement, strip=False):
        pr = evolution.gh_kerr_schild_shell_problem(refinement, N)
        ids = np.arange(pr.brick.n_elements)
        x, J, stat = pr.coords(ids), pr.inverse_jacobian(ids), pr.static(ids)
        u0 = pr.u0(ids, 0.0)
        H = np.zeros((len(x), 4, N ** 3))
        dH = np.zeros((len(x), 16, N ** 3))
        for e in range(len(x)):
            H[e], dH[e] = orc.analytic_christoffel_gauge(N, u0[e], J[e])
        nd, perm = pr.orientations
        mt = np.array(pr.mortars).copy()
        if strip:
            mt[:, 3] &= 7
        r = orc.dg_rhs(1, N, u0, J, np.concatenate([stat, H, dH], axis=1), pr.neighbors,
                       gauge_params=orc.GAUGE_GIVEN, nbr_dir=nd, face_perm=perm,
                       mortars=mt if len(mt) else None)
        return float(np.max(np.abs(r)))
    coarse = max_rhs((0, 0))
    for w in (0, 5):
        ref = [(0, 0)] * 6
        ref[w] = (1, 1)
        assert max_rhs([ref]) < 1.5 * coarse
        assert max_rhs([ref], strip=True) > 20 * coarse
