"""BinaryCompactObject domain (BASELINE configs[4]; host-side set-up):
the block maps are restated from the reference's Wedge / Frustum forward maps
(Wedge.cpp:250-537, Frustum.cpp:30-213) and assembled like BinaryCompactObject.cpp:420-552.
The reference's own tests check such a creator by its block structure (44 blocks, each
internal face shared by exactly two blocks, the external faces on the two excision spheres
and the outer sphere); the same checks here, plus the coincidence of the grid points on every
internal face, which the conforming face kernels rely on."""
import numpy as np
import pytest

from spectre_b200 import bco, domain


@pytest.fixture(scope="module")
def dom():
    return bco.BinaryCompactObject(8.0, -8.0, 0.8, 4.0, 1.1, 3.5, 60.0, 300.0, 0, 4,
                                   opening_angle_degrees=120.0)


def test_block_structure(dom):
    assert dom.n_blocks == 44 and dom.n_elements == 44
    nbr = dom.neighbors()
    ext = [(e, d) for e in range(44) for d in range(6) if nbr[e, d] == -1]
    assert len(ext) == 22 and len(dom.mortars()) == 0
    kinds = [dom.external_boundary(e, d) for e, d in ext]
    assert kinds.count("excision_a") == 6 and kinds.count("excision_b") == 6
    assert kinds.count("outer") == 10
    # external faces lie on the three spheres
    X = dom.coords()
    fp = [domain._face_point_indices(4, d) for d in range(6)]
    centre = {"excision_a": (8.0, 0, 0), "excision_b": (-8.0, 0, 0), "outer": (0, 0, 0)}
    radius = {"excision_a": 0.8, "excision_b": 1.1, "outer": 300.0}
    for (e, d), kind in zip(ext, kinds):
        r = np.linalg.norm(X[e][:, fp[d]] - np.array(centre[kind])[:, None], axis=0)
        assert np.max(np.abs(r - radius[kind])) < 1e-12 * radius[kind]
    # every internal face is shared by exactly two elements, symmetric table
    nd, perm = dom.neighbor_orientations()
    for e in range(44):
        for d in range(6):
            if nbr[e, d] >= 0:
                assert nbr[nbr[e, d], nd[e, d]] == e and nd[nbr[e, d], nd[e, d]] == d


def test_grid_points_coincide_on_internal_faces(dom):
    N = 4
    nbr = dom.neighbors()
    nd, perm = dom.neighbor_orientations()
    X = dom.coords()
    fp = [domain._face_point_indices(N, d) for d in range(6)]
    q = np.arange(N * N)
    qa, qb = q % N, q // N
    worst = 0.0
    for e in range(dom.n_elements):
        for d in range(6):
            e2 = nbr[e, d]
            if e2 < 0:
                continue
            code = perm[e, d]
            na, nb = (qb, qa) if code & 1 else (qa, qb)
            if code & 2:
                na = N - 1 - na
            if code & 4:
                nb = N - 1 - nb
            mine = X[e][:, fp[d]]
            theirs = X[e2][:, fp[nd[e, d]]][:, na + N * nb]
            worst = max(worst, np.max(np.abs(mine - theirs)) / np.abs(mine).max())
    assert worst < 1e-12


def test_complex_step_jacobian_and_orientation(dom):
    """the complex-step Jacobian agrees with central differences of the forward map, all
    blocks are right-handed, and the block volumes add up to the volume between the spheres"""
    xi = np.array([[0.3], [-0.2], [0.55]])
    for e in (0, 7, 13, 20, 24, 27, 32, 33, 36, 43):
        x, jac = dom.map_points(e, xi)
        h = 1e-6
        for j in range(3):
            dp, dm = xi.copy(), xi.copy()
            dp[j] += h
            dm[j] -= h
            fd = (dom.map_points(e, dp)[0] - dom.map_points(e, dm)[0]) / (2 * h)
            assert np.max(np.abs(fd[:, 0] - jac[:, j, 0])) < 1e-7 * max(1.0, np.abs(jac).max())
        assert np.linalg.det(jac[:, :, 0]) > 0.0
    fine = bco.BinaryCompactObject(8.0, -8.0, 0.8, 4.0, 1.1, 3.5, 60.0, 300.0, 0, 12,
                                   opening_angle_degrees=120.0)
    w = fine.weights
    w3 = (w[:, None, None] * w[None, :, None] * w[None, None, :]).transpose(2, 1, 0).ravel()
    vol = 0.0
    for e in range(fine.n_elements):
        Jinv = fine.inverse_jacobian([e])[0]
        M = np.array([[Jinv[jh + 3 * i] for i in range(3)] for jh in range(3)])
        det = 1.0 / np.linalg.det(np.moveaxis(M, -1, 0))
        vol += np.sum(w3 * det)
    exact = 4.0 / 3.0 * np.pi * (300.0 ** 3 - 0.8 ** 3 - 1.1 ** 3)
    assert abs(vol - exact) < 2e-3 * exact


def test_refinement_groups_make_hanging_faces():
    """per-group refinement as in Inspiral.yaml:95-101 (cubes one level finer in the angular
    directions): the interfaces between the groups become 2:1 mortars"""
    ref = {g: (0, 0, 0) for g in bco.BinaryCompactObject.GROUPS}
    ref["ObjectACube"] = ref["ObjectBCube"] = (1, 1, 0)
    d = bco.BinaryCompactObject(8.0, -8.0, 0.8, 4.0, 0.8, 4.0, 60.0, 300.0, ref, 3)
    assert d.n_elements == 32 + 12 * 4
    m = d.mortars()
    nbr = d.neighbors()
    assert len(m) > 0 and (nbr == domain.HANGING).sum() > 0
    # four mortars on each of the 12 faces towards the shells and the 10 faces towards the
    # envelope (the two cube faces that meet between the objects are equally fine)
    assert len(m) == (12 + 10) * 4
    assert ((nbr == -1).sum()) == 22
