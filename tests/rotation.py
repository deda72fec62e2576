"""Test helper: give every element of a mesh of aligned blocks its own rotated /
reflected logical frame (one of the 48 orientations of a cube) and build the
tables a mesh of non-aligned blocks needs — neighbour directions, face
permutations (OrientationMap restricted to the face, as orient_variables_on_slice
applies it: Domain/Structure/OrientationMapHelpers.cpp:25-120) and mortar rows.
The physics must not notice: results mapped back equal those of the aligned mesh."""
import itertools

import numpy as np


def signed_perms():
    out = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            out.append((perm, signs))
    return out  # the 48 orientations of a cube


def point_map(N, perm, signs):
    """new_index[p_old] for the frame xi'_a = signs[a] * xi_{perm[a]}."""
    p = np.arange(N ** 3)
    old = (p % N, (p // N) % N, p // (N * N))
    new = []
    for a in range(3):
        i = old[perm[a]]
        new.append(i if signs[a] > 0 else N - 1 - i)
    return new[0] + N * (new[1] + N * new[2])


def dir_map(perm, signs):
    """old direction -> new direction."""
    m = {}
    for a in range(3):
        for side in range(2):
            old_d = 2 * perm[a] + (side if signs[a] > 0 else 1 - side)
            m[old_d] = 2 * a + side
    return m


def face_points(N, d):
    dim, fixed = d // 2, (N - 1 if d % 2 else 0)
    q = np.arange(N * N)
    a, b = q % N, q // N
    return [fixed + N * (a + N * b), a + N * (fixed + N * b), a + N * (b + N * fixed)][dim]


def _face_code(N, pm_e, d_new, d_old, pm_nb, nd_new):
    """Permutation code taking face point (a, b) of our rotated face d_new to the
    point of the neighbour's rotated face nd_new that it touches (the neighbour
    sits across old direction d_old, with the same tangential indices there)."""
    inv = np.argsort(pm_e)  # new index -> old index
    p_old = inv[face_points(N, d_new)]
    i = [p_old % N, (p_old // N) % N, p_old // (N * N)]
    dim = d_old // 2
    i[dim] = np.where(i[dim] == 0, N - 1, 0)
    p_nb_new = pm_nb[i[0] + N * (i[1] + N * i[2])]
    pos = {int(v): k for k, v in enumerate(face_points(N, nd_new))}
    target = np.array([pos[int(v)] for v in p_nb_new])
    q = np.arange(N * N)
    qa, qb = q % N, q // N
    for code in range(8):
        na = np.where(code & 1, qb, qa)
        nbb = np.where(code & 1, qa, qb)
        if code & 2:
            na = N - 1 - na
        if code & 4:
            nbb = N - 1 - nbb
        if np.array_equal(na + N * nbb, target):
            return code
    raise AssertionError("no face permutation matches")


def rotate_problem(N, u, J, stat, nbr, frames, mortars=None):
    """Returns u, J, stat, nbr, neighbour directions, face permutations in the rotated
    frames, the per-element point maps (new_index[p_old]) and, if given aligned
    mortar rows, the rows of the rotated mesh (fine direction | perm << 3; mortar
    sizes along the rotated coarse face's dimensions)."""
    ne = u.shape[0]
    pm = [point_map(N, *frames[e]) for e in range(ne)]
    dm = [dir_map(*frames[e]) for e in range(ne)]
    u_r, J_r, s_r = np.empty_like(u), np.empty_like(J), np.empty_like(stat)
    nbr_r = np.full_like(nbr, -1)
    nd_r = np.zeros_like(nbr)
    perm_r = np.zeros_like(nbr)
    for e in range(ne):
        perm, signs = frames[e]
        u_r[e][:, pm[e]] = u[e]
        s_r[e][:, pm[e]] = stat[e]
        for a in range(3):
            for i in range(3):
                J_r[e][a + 3 * i][pm[e]] = signs[a] * J[e][perm[a] + 3 * i]
    for e in range(ne):
        for d_old in range(6):
            nb = nbr[e, d_old]
            d_new = dm[e][d_old]
            nbr_r[e, d_new] = nb
            if nb < 0:          # external, ghost slot or a sentinel: nothing to orient
                nd_r[e, d_new] = d_new ^ 1
                continue
            nd_new = dm[nb][d_old ^ 1]
            nd_r[e, d_new] = nd_new
            perm_r[e, d_new] = _face_code(N, pm[e], d_new, d_old, pm[nb], nd_new)
    if mortars is None:
        return u_r, J_r, s_r, nbr_r, nd_r, perm_r, pm
    rows = []
    for ec, dc, ef, df, sa, sb in np.asarray(mortars).tolist():
        assert df == (dc ^ 1), "input rows must be those of aligned blocks"
        perm, signs = frames[ec]
        dc_new, df_new = dm[ec][dc], dm[ef][df]
        old_tan = [t for t in range(3) if t != dc // 2]
        size_old = {old_tan[0]: sa, old_tan[1]: sb}
        sizes = []
        for t in [t for t in range(3) if t != dc_new // 2]:     # rotated face dimensions
            s = size_old[perm[t]]
            if signs[t] < 0 and s != 0:
                s = 3 - s                                        # LowerHalf <-> UpperHalf
            sizes.append(s)
        code = _face_code(N, pm[ec], dc_new, dc, pm[ef], df_new)
        rows.append([ec, dc_new, ef, df_new | (code << 3), sizes[0], sizes[1]])
    return u_r, J_r, s_r, nbr_r, nd_r, perm_r, pm, np.array(rows, dtype=np.int32)
