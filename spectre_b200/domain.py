"""Host-side domain data model for the DG path: a periodic/non-periodic Brick
of 2^L elements per dimension and the spherical shell of six Wedge blocks per
radial layer (Domain/Creators/Sphere.cpp with an excised interior), elements
ordered along the per-block Z-curve
(Morton order) like the reference places them (Domain/Structure/ZCurve.cpp:
17-80, DgElementArray.hpp:53-66), and the contiguous equal-cost partition of
that order over ranks (ElementDistribution.hpp:33-47 with NumGridPoints
weights, which is uniform for an isotropic mesh).

What leaves this module is exactly what the reference's initialisation puts in
the DataBox for this path: inertial coordinates, the inverse
Jacobian of the affine map (CoordinateMaps/Affine.cpp), and the neighbour
table of every element (Domain/Creators/Rectilinear.cpp with Periodic boundary
conditions: the block is its own neighbour with aligned orientation).
"""
from __future__ import annotations

import numpy as np

from . import lib


def element_id(block: int, segment_indices, refinement_levels, grid_index: int = 0) -> int:
    """64-bit ElementId packing (Domain/Structure/ElementId.hpp:29-110): block
    8 bits, grid index 4 bits, direction 4 bits (unused here: 0), then per
    dimension a 12-bit index and a 4-bit refinement level."""
    v = (block & 0xFF) | ((grid_index & 0xF) << 8)
    shift = 16
    for idx, lev in zip(segment_indices, refinement_levels):
        v |= (idx & 0xFFF) << shift
        v |= (lev & 0xF) << (shift + 12)
        shift += 16
    return v


def z_curve_index(ix: int, iy: int, iz: int, levels) -> int:
    """Domain/Structure/ZCurve.cpp:17-80 for one block."""
    dims = sorted([(levels[0], 0), (levels[1], 1), (levels[2], 2)])
    idx = (ix, iy, iz)
    out = 0
    leading_gap = 0
    for i in range(3):
        lev, dim = dims[i]
        total_gap = leading_gap
        if lev > 0:
            leading_gap += 1
        for bit in range(lev):
            out |= (idx[dim] & (1 << bit)) << total_gap
            for j in range(3):
                if i != j and bit + 1 < dims[j][0]:
                    total_gap += 1
    return out


class Brick:
    """DomainCreator Brick (Domain/Creators/Rectilinear.cpp), one block."""

    def __init__(self, lower, upper, refinement, N, periodic=(True, True, True),
                 order="zcurve"):
        self.lower = np.asarray(lower, float)
        self.upper = np.asarray(upper, float)
        self.levels = tuple(int(r) for r in refinement)
        self.ne = tuple(2 ** r for r in self.levels)
        self.N = int(N)
        self.n = self.N ** 3
        self.periodic = tuple(periodic) if not isinstance(periodic, bool) else (periodic,) * 3
        nx, ny, nz = self.ne
        cells = [(ix, iy, iz) for iz in range(nz) for iy in range(ny) for ix in range(nx)]
        if order == "zcurve":
            cells.sort(key=lambda c: z_curve_index(c[0], c[1], c[2], self.levels))
        self.cells = cells  # position in this list = element index
        self.index_of = {c: i for i, c in enumerate(cells)}
        self.n_elements = len(cells)
        self.xi, self.weights = lib.collocation_points_and_weights(self.N)

    def element_ids(self):
        return [element_id(0, c, self.levels) for c in self.cells]

    def _bounds(self, cell):
        h = (self.upper - self.lower) / np.asarray(self.ne)
        lo = self.lower + h * np.asarray(cell)
        return lo, lo + h

    def coords(self, ids=None):
        """Inertial coordinates [len(ids), 3, n] (all elements if ids is None)."""
        N, n = self.N, self.n
        ids = range(self.n_elements) if ids is None else ids
        p = np.arange(n)
        idx = (p % N, (p // N) % N, p // (N * N))
        out = np.zeros((len(ids), 3, n))
        for k, e in enumerate(ids):
            lo, hi = self._bounds(self.cells[e])
            for d in range(3):
                out[k, d] = 0.5 * (hi[d] - lo[d]) * self.xi[idx[d]] + 0.5 * (hi[d] + lo[d])
        return out

    def inverse_jacobian(self, ids=None):
        ids = range(self.n_elements) if ids is None else ids
        out = np.zeros((len(ids), 9, self.n))
        for k, e in enumerate(ids):
            lo, hi = self._bounds(self.cells[e])
            for d in range(3):
                out[k, d + 3 * d] = 2.0 / (hi[d] - lo[d])
        return out

    def neighbors(self):
        """[n_elements, 6] global element index or -1 (external boundary)."""
        nb = np.full((self.n_elements, 6), -1, dtype=np.int32)
        for e, cell in enumerate(self.cells):
            for d in range(6):
                dim, side = d // 2, d % 2
                c = list(cell)
                c[dim] += 1 if side else -1
                if c[dim] < 0 or c[dim] >= self.ne[dim]:
                    if not self.periodic[dim]:
                        continue
                    c[dim] %= self.ne[dim]
                nb[e, d] = self.index_of[tuple(c)]
        return nb


HANGING = -2 ** 31   # neighbour-table entry of a face handled by the mortar table
MORTAR_FULL, MORTAR_LOWER_HALF, MORTAR_UPPER_HALF = 0, 1, 2


class RefinedBrick:
    """A Brick whose coarse cells (2^L per dimension) are individually h-refined
    once, isotropically (2 x 2 x 2 children) or only in some dimensions (the
    reference's AMR is anisotropic: `Amr: Policies: Isotropy: Anisotropic`), as AMR
    or per-block `InitialRefinement` differences produce: element faces between a
    coarser and a finer element are non-conforming 2:1 mortars (dg::mortar_size,
    NumericalAlgorithms/DiscontinuousGalerkin/MortarHelpers.cpp:51-77;
    Element<3>::neighbors() then holds several neighbour ids for that direction,
    Domain/Structure/Neighbors.hpp).

    refined_cells: cells split in all three dimensions, or a dict cell -> (bool,
    bool, bool) saying which dimensions are split.
    neighbors(): [n_elements, 6] with HANGING on both sides of such faces;
    mortars(): rows (coarse element, its direction, fine element, its direction,
    size_a, size_b), the MortarSize of the fine face inside the coarse face per
    face dimension (first remaining dimension first; Full where both sides have
    the same extent).  An interface where each side is the finer one in one face
    dimension (the mortar would be smaller than both faces) is not supported."""

    def __init__(self, lower, upper, refinement, N, refined_cells, periodic=(True, True, True)):
        self.lower = np.asarray(lower, float)
        self.upper = np.asarray(upper, float)
        self.levels = tuple(int(r) for r in refinement)
        self.ne = tuple(2 ** r for r in self.levels)
        self.N = int(N)
        self.n = self.N ** 3
        self.periodic = tuple(periodic) if not isinstance(periodic, bool) else (periodic,) * 3
        if isinstance(refined_cells, dict):
            self.split = {tuple(c): tuple(bool(x) for x in m) for c, m in refined_cells.items()}
        else:
            self.split = {tuple(c): (True, True, True) for c in refined_cells}
        self.refined = set(self.split)
        nx, ny, nz = self.ne
        cells = [(ix, iy, iz) for iz in range(nz) for iy in range(ny) for ix in range(nx)]
        cells.sort(key=lambda c: z_curve_index(c[0], c[1], c[2], self.levels))
        # element = (coarse cell, child or None); children in Z order, child
        # coordinate 0 in a dimension that is not split
        self.elements = []
        for c in cells:
            m = self.split.get(c)
            if m is None:
                self.elements.append((c, None))
                continue
            for k in range(8):
                ch = (k & 1, (k >> 1) & 1, (k >> 2) & 1)
                if all(ch[d] == 0 or m[d] for d in range(3)):
                    self.elements.append((c, ch))
        self.index_of = {el: i for i, el in enumerate(self.elements)}
        self.n_elements = len(self.elements)
        self.xi, self.weights = lib.collocation_points_and_weights(self.N)
        self._tables = None

    def _mask(self, c):
        return self.split.get(c, (False, False, False))

    def element_ids(self):
        out = []
        for c, ch in self.elements:
            m = self._mask(c)
            chh = ch or (0, 0, 0)
            out.append(element_id(0, [2 * c[d] + chh[d] if m[d] else c[d] for d in range(3)],
                                  [self.levels[d] + (1 if m[d] else 0) for d in range(3)]))
        return out

    def _bounds(self, el):
        c, ch = el
        h = (self.upper - self.lower) / np.asarray(self.ne)
        lo = self.lower + h * np.asarray(c)
        if ch is not None:
            m = np.asarray(self._mask(c), float)
            h = h * (1.0 - 0.5 * m)
            lo = lo + h * np.asarray(ch) * m
        return lo, lo + h

    def coords(self, ids=None):
        N, n = self.N, self.n
        ids = range(self.n_elements) if ids is None else ids
        p = np.arange(n)
        idx = (p % N, (p // N) % N, p // (N * N))
        out = np.zeros((len(ids), 3, n))
        for k, e in enumerate(ids):
            lo, hi = self._bounds(self.elements[e])
            for d in range(3):
                out[k, d] = 0.5 * (hi[d] - lo[d]) * self.xi[idx[d]] + 0.5 * (hi[d] + lo[d])
        return out

    def inverse_jacobian(self, ids=None):
        ids = range(self.n_elements) if ids is None else ids
        out = np.zeros((len(ids), 9, self.n))
        for k, e in enumerate(ids):
            lo, hi = self._bounds(self.elements[e])
            for d in range(3):
                out[k, d + 3 * d] = 2.0 / (hi[d] - lo[d])
        return out

    def _build(self):
        nb = np.full((self.n_elements, 6), -1, dtype=np.int64)
        mortars = []
        for e, (c, ch) in enumerate(self.elements):
            m = self._mask(c)
            chh = ch or (0, 0, 0)
            for d in range(6):
                dim, side = d // 2, d % 2
                if m[dim] and chh[dim] != side:
                    sib = list(chh)          # sibling inside the same refined cell
                    sib[dim] = side
                    nb[e, d] = self.index_of[(c, tuple(sib))]
                    continue
                nc = list(c)
                nc[dim] += 1 if side else -1
                if nc[dim] < 0 or nc[dim] >= self.ne[dim]:
                    if not self.periodic[dim]:
                        continue
                    nc[dim] %= self.ne[dim]
                nc = tuple(nc)
                mn = self._mask(nc)
                fd = [x for x in range(3) if x != dim]
                # per face dimension: +1 the neighbour is finer, -1 we are finer, 0 equal
                rel = [int(mn[x]) - int(m[x]) for x in fd]
                if 1 in rel and -1 in rel:
                    raise NotImplementedError("mortar smaller than both faces")
                base = [0, 0, 0]
                if mn[dim]:
                    base[dim] = 1 - side
                if all(r <= 0 for r in rel):
                    # one neighbour: the same size or coarser
                    for x in fd:
                        base[x] = chh[x] if mn[x] else 0
                    other = (nc, tuple(base)) if nc in self.split else (nc, None)
                    if any(r < 0 for r in rel):
                        nb[e, d] = HANGING   # the coarser side lists the mortar
                    else:
                        nb[e, d] = self.index_of[other]
                    continue
                # we are the coarse side: one mortar per finer neighbour
                nb[e, d] = HANGING
                ra = range(2) if rel[0] == 1 else [None]
                rb = range(2) if rel[1] == 1 else [None]
                for kb in rb:
                    for ka in ra:
                        child = list(base)
                        child[fd[0]] = ka if ka is not None else (chh[fd[0]] if mn[fd[0]] else 0)
                        child[fd[1]] = kb if kb is not None else (chh[fd[1]] if mn[fd[1]] else 0)
                        size = lambda k: MORTAR_FULL if k is None else (
                            MORTAR_UPPER_HALF if k else MORTAR_LOWER_HALF)
                        mortars.append((e, d, self.index_of[(nc, tuple(child))], d ^ 1,
                                        size(ka), size(kb)))
        self._tables = (nb.astype(np.int32), np.asarray(mortars, dtype=np.int32).reshape(-1, 6))

    def neighbors(self):
        if self._tables is None:
            self._build()
        return self._tables[0]

    def mortars(self):
        if self._tables is None:
            self._build()
        return self._tables[1]


# ---------------------------------------------------------------------------
# Wedge<3> (Domain/CoordinateMaps/Wedge.cpp:67-130 constructor constants,
# :251-291 cap functions, :327-360 1/rho, :362-435 radial functions, :476-537
# forward map) for a centred wedge with spherical inner and outer surfaces
# (sphericity 1, no focal offset), and the six orientations of
# orientations_for_sphere_wrappings (Domain/DomainHelpers.cpp:553-578).
# ---------------------------------------------------------------------------
# WEDGE_ORIENTATIONS[w][i] = (source dimension, sign): discrete_rotation
# (Domain/Structure/OrientationMap.cpp:219-234) sets
# new_coords[i] = sign * source_coords[dimension].
WEDGE_ORIENTATIONS = (
    ((0, 1), (1, 1), (2, 1)),     # upper z: aligned
    ((0, 1), (1, -1), (2, -1)),   # lower z: (upper_xi, lower_eta, lower_zeta)
    ((0, 1), (2, 1), (1, -1)),    # upper y: (upper_xi, upper_zeta, lower_eta)
    ((0, 1), (2, -1), (1, 1)),    # lower y: (upper_xi, lower_zeta, upper_eta)
    ((2, 1), (0, 1), (1, 1)),     # upper x: (upper_zeta, upper_xi, upper_eta)
    ((2, -1), (0, -1), (1, 1)),   # lower x: (lower_zeta, lower_xi, upper_eta)
)


def wedge_map(xi, eta, zeta, r_in, r_out, wedge, equiangular=True, distribution="Linear"):
    """Block-logical coordinates in [-1,1]^3 -> (x [3,...], jacobian [3,3,...])
    with jacobian[i][j] = d x^i / d xi^j.  wedge: index into WEDGE_ORIENTATIONS
    or an orientation ((dim, sign) x 3) itself."""
    orientation = WEDGE_ORIENTATIONS[wedge] if isinstance(wedge, int) else wedge
    xi, eta, zeta = (np.asarray(v, float) for v in (xi, eta, zeta))
    if equiangular:
        cap = [np.tan(0.25 * np.pi * xi), np.tan(0.25 * np.pi * eta)]
        dcap = [0.25 * np.pi / np.cos(0.25 * np.pi * xi) ** 2,
                0.25 * np.pi / np.cos(0.25 * np.pi * eta) ** 2]
    else:
        cap = [xi, eta]
        dcap = [np.ones_like(xi), np.ones_like(eta)]
    one_over_rho = 1.0 / np.sqrt(1.0 + cap[0] ** 2 + cap[1] ** 2)
    if distribution == "Linear":
        s = 0.5 * (r_out + r_in) + 0.5 * (r_out - r_in) * zeta
        ds = 0.5 * (r_out - r_in) * np.ones_like(zeta)
    elif distribution == "Logarithmic":
        s = np.exp(0.5 * np.log(r_out * r_in) + 0.5 * np.log(r_out / r_in) * zeta)
        ds = 0.5 * s * np.log(r_out / r_in)
    elif distribution == "Inverse":
        s = 2.0 / ((1.0 + zeta) / r_out + (1.0 - zeta) / r_in)
        ds = 2.0 * (r_in * r_out ** 2 - r_in ** 2 * r_out) / \
            (r_in + r_out + zeta * (r_in - r_out)) ** 2
    else:
        raise ValueError(f"unsupported radial distribution {distribution}")
    z = s * one_over_rho                     # generalized z
    # source frame: (polar, azimuth, radial) = z * (cap0, cap1, 1)
    dz = [-s * one_over_rho ** 3 * cap[0] * dcap[0],
          -s * one_over_rho ** 3 * cap[1] * dcap[1],
          ds * one_over_rho]
    src = [z * cap[0], z * cap[1], z]
    dsrc = [[z * dcap[0] + cap[0] * dz[0], cap[0] * dz[1], cap[0] * dz[2]],
            [cap[1] * dz[0], z * dcap[1] + cap[1] * dz[1], cap[1] * dz[2]],
            [dz[0], dz[1], dz[2]]]
    x, jac = [], []
    for i in range(3):
        dim, sign = orientation[i]
        x.append(sign * src[dim])
        jac.append([sign * dsrc[dim][j] for j in range(3)])
    return np.array(x), np.array(jac)


def _face_point_indices(N, d):
    """Volume indices of the face points of direction d = 2 dim + side, in the
    face's own ordering (first remaining dimension fastest)."""
    dim, fixed = d // 2, (N - 1 if d % 2 else 0)
    q = np.arange(N * N)
    a, b = q % N, q // N
    return [fixed + N * (a + N * b), a + N * (fixed + N * b), a + N * (b + N * fixed)][dim]


def connectivity_from_geometry(coords, N, tol=1e-9):
    """Neighbour table of a conforming multi-block mesh from the coincidence of
    element faces: returns (neighbors, neighbor_direction, face_permutation),
    each [n_elements, 6].  neighbors = -1 on external faces.  The permutation
    code is the OrientationMap between the two blocks restricted to the face
    (Domain/Structure/OrientationMapHelpers.cpp:25-120): bit 0 = the two face
    coordinates are swapped, bit 1 / bit 2 = the neighbour's first / second face
    coordinate runs backwards."""
    from scipy.spatial import cKDTree
    ne = coords.shape[0]
    fp = [_face_point_indices(N, d) for d in range(6)]
    centers = np.empty((ne * 6, 3))
    for d in range(6):
        centers[d::6] = coords[:, :, fp[d]].mean(axis=2)
    # coincident faces come in pairs: the nearest other face centre is the partner
    # if it is closer than tol relative to the local length scale (the domain may
    # span many orders of magnitude in radius)
    tree = cKDTree(centers)
    dist, idx = tree.query(centers, k=2)
    me = np.arange(len(centers))
    partner = np.where(idx[:, 0] == me, idx[:, 1], idx[:, 0])   # (self may come second)
    local_scale = np.maximum(np.linalg.norm(centers, axis=1), 1e-300)
    is_pair = dist[:, 1] < tol * local_scale
    pairs = [(int(i), int(partner[i])) for i in np.nonzero(is_pair)[0] if i < partner[i]]
    nbr = np.full((ne, 6), -1, dtype=np.int32)
    nd = np.tile((np.arange(6) ^ 1).astype(np.int32), (ne, 1))
    perm = np.zeros((ne, 6), dtype=np.int32)
    q = np.arange(N * N)
    qa, qb = q % N, q // N
    targets = []
    for code in range(8):
        na = np.where(code & 1, qb, qa)
        nb = np.where(code & 1, qa, qb)
        if code & 2:
            na = N - 1 - na
        if code & 4:
            nb = N - 1 - nb
        targets.append(na + N * nb)

    def match(e, d, e2, d2):
        mine = coords[e][:, fp[d]]
        theirs = coords[e2][:, fp[d2]]
        scale = np.abs(mine).max()
        for code in range(8):
            if np.max(np.abs(mine - theirs[:, targets[code]])) < tol * scale:
                return code
        raise ValueError(f"faces ({e},{d}) and ({e2},{d2}) coincide but their points do not")

    for i, j in pairs:
        e, d, e2, d2 = i // 6, i % 6, j // 6, j % 6
        if partner[j] != i or not is_pair[j]:
            raise ValueError("a face has more than one neighbour (non-conforming mesh)")
        nbr[e, d], nd[e, d], perm[e, d] = e2, d2, match(e, d, e2, d2)
        nbr[e2, d2], nd[e2, d2], perm[e2, d2] = e, d, match(e2, d2, e, d)
    return nbr, nd, perm


class SphericalShell:
    """DomainCreator Sphere with an excised interior (Domain/Creators/Sphere.cpp,
    `Interior: ExciseWithBoundaryCondition`): per radial layer six Wedge<3> blocks
    (upper/lower z, y, x in the order of sph_wedge_coordinate_maps,
    Domain/DomainHelpers.cpp:596-700), layers ordered inside-out, every block of a
    layer refined to 2^L_angular x 2^L_angular x 2^L_radial elements.  The
    refinement may differ from layer to layer by one angular level (the creator's
    per-block `InitialRefinement`): the spherical interface between such layers
    is made of non-conforming 2:1 mortars (neighbors() = HANGING, mortars())."""

    def __init__(self, inner_radius, outer_radius, refinement, N, radial_partitioning=(),
                 radial_distribution="Logarithmic", equiangular=True, order="block"):
        """refinement: int, (angular, radial) initial refinement levels, or a list of
        (angular, radial) per layer.
        order: "block" = block-major, Z-curve inside a block (the reference's element
        placement, ElementDistribution.hpp:33-47); "radial" = spherical layers of
        elements inside-out (all six wedges of a radial index together, Z-curve in
        the angular directions), so that a contiguous partition cuts the shell at
        constant radius and the halo is 2 x 6 x 4^L faces per rank."""
        self.N = int(N)
        self.n = self.N ** 3
        self.radii = [float(inner_radius), *map(float, radial_partitioning), float(outer_radius)]
        self.n_layers = len(self.radii) - 1
        if isinstance(refinement, int):
            refinement = (refinement, refinement)
        if not isinstance(refinement[0], (tuple, list)):
            refinement = [tuple(refinement)] * self.n_layers
        assert len(refinement) == self.n_layers
        # a layer entry is (angular, radial) for its six wedges or six such pairs, one
        # per wedge (the creator's per-block InitialRefinement)
        self.refinement = [
            [(int(w[0]), int(w[1])) for w in r] if isinstance(r[0], (tuple, list))
            else (int(r[0]), int(r[1])) for r in refinement]
        self.block_levels = []
        for r in self.refinement:
            per_wedge = r if isinstance(r, list) else [r] * 6
            assert len(per_wedge) == 6
            self.block_levels += [(w[0], w[0], w[1]) for w in per_wedge]
        self.levels = self.block_levels[0]
        self.ne = tuple(2 ** r for r in self.levels)
        if isinstance(radial_distribution, str):
            radial_distribution = [radial_distribution] * self.n_layers
        self.distributions = list(radial_distribution)
        self.equiangular = bool(equiangular)
        self.n_blocks = 6 * self.n_layers
        self.cells = []  # element index -> (block, cell)
        radial_offset, off = [], 0
        for layer in range(self.n_layers):
            for w in range(6):
                lev = self.block_levels[6 * layer + w]
                nx, ny, nz = (2 ** l for l in lev)
                cells = [(ix, iy, iz) for iz in range(nz) for iy in range(ny) for ix in range(nx)]
                cells.sort(key=lambda c: z_curve_index(c[0], c[1], c[2], lev))
                self.cells += [(6 * layer + w, c) for c in cells]
            radial_offset.append(off)
            off += max(2 ** self.block_levels[6 * layer + w][2] for w in range(6))
        if order == "radial":
            if any(len(set(self.block_levels[6 * l:6 * l + 6])) > 1 for l in range(self.n_layers)):
                raise ValueError('order="radial" needs the same refinement in the six wedges '
                                 'of a layer')
            self.cells.sort(key=lambda bc: (
                radial_offset[bc[0] // 6] + bc[1][2], bc[0] % 6,
                z_curve_index(bc[1][0], bc[1][1], 0,
                              (self.block_levels[bc[0]][0], self.block_levels[bc[0]][1], 0))))
        elif order != "block":
            raise ValueError(order)
        self.order = order
        self.n_elements = len(self.cells)
        self.xi, self.weights = lib.collocation_points_and_weights(self.N)
        self._conn = None

    def element_ids(self):
        return [element_id(b, c, self.block_levels[b]) for b, c in self.cells]

    def map_points(self, e, xi):
        """Element-logical points xi [3, m] of element e -> (x [3, m], jacobian
        [3, 3, m] = d x^i / d xi^j with respect to the ELEMENT-logical coordinates)."""
        b, cell = self.cells[e]
        layer, wedge = b // 6, b % 6
        blk, half = [], []
        for d in range(3):
            h = 2.0 / 2 ** self.block_levels[b][d]
            blk.append(-1.0 + h * cell[d] + 0.5 * h * (np.asarray(xi[d], float) + 1.0))
            half.append(0.5 * h)
        x, jac = wedge_map(blk[0], blk[1], blk[2], self.radii[layer], self.radii[layer + 1],
                           wedge, self.equiangular, self.distributions[layer])
        for j in range(3):
            jac[:, j] *= half[j]
        return x, jac

    def _geometry(self, ids):
        N, n = self.N, self.n
        p = np.arange(n)
        idx = (p % N, (p // N) % N, p // (N * N))
        xi = [self.xi[idx[d]] for d in range(3)]
        X = np.empty((len(ids), 3, n))
        Jinv = np.empty((len(ids), 9, n))
        for k, e in enumerate(ids):
            x, jac = self.map_points(e, xi)
            X[k] = x
            inv = np.linalg.inv(np.moveaxis(jac, -1, 0))   # [n, jhat, i] = d xi^jhat / d x^i
            for jh in range(3):
                for i in range(3):
                    Jinv[k, jh + 3 * i] = inv[:, jh, i]
        return X, Jinv

    def coords(self, ids=None):
        ids = range(self.n_elements) if ids is None else ids
        return self._geometry(list(ids))[0]

    def inverse_jacobian(self, ids=None):
        ids = range(self.n_elements) if ids is None else ids
        return self._geometry(list(ids))[1]

    def _connectivity(self):
        # the topology does not depend on the polynomial order: match the faces of
        # the 2-point (corner) mesh, whose 2 x 2 face points still tell the eight
        # face permutations apart
        if self._conn is None:
            corners = SphericalShell(self.radii[0], self.radii[-1], self.refinement, 2,
                                     self.radii[1:-1], self.distributions, self.equiangular,
                                     self.order)
            nbr, nd, perm = connectivity_from_geometry(corners.coords(), 2)
            mortars = find_hanging_faces(corners, nbr)
            # what is still unmatched must be the two spheres; an angular face left over is
            # an interface this path cannot represent (levels differing by more than one, or
            # a mortar smaller than both faces: finer in one face dimension, coarser in the other)
            bad = [(e, d) for e, d in zip(*np.nonzero(nbr == -1)) if d < 4]
            if bad:
                e, d = bad[0]
                raise ValueError(f"unsupported non-conforming interface at block {self.cells[e][0]} "
                                 f"direction {d}: neighbouring blocks may differ by one refinement "
                                 "level, and not in opposite senses in the two face dimensions")
            self._conn = (nbr, nd, perm, mortars)
        return self._conn

    def neighbors(self):
        return self._connectivity()[0]

    def neighbor_orientations(self):
        """(neighbor_direction, face_permutation), see connectivity_from_geometry."""
        c = self._connectivity()
        return c[1], c[2]

    def mortars(self):
        """Non-conforming mortars between layers of different angular refinement,
        rows (coarse element, direction, fine element, direction, size_a, size_b)."""
        return self._connectivity()[3]

    def external_boundary(self, e, d):
        """'inner' (excision) or 'outer' for an external face."""
        b, cell = self.cells[e]
        assert d // 2 == 2
        return "inner" if d == 4 else "outer"


def find_hanging_faces(dom, nbr, tol=1e-9):
    """2:1 non-conforming faces of a multi-block domain with `map_points`: a face
    without a conforming partner whose logical halves / quarters (split in one or
    both face dimensions) have the centres of (likewise unmatched) faces at their
    centres is a coarse face with two or four mortars.  Marks both sides HANGING in
    nbr (in place) and returns the mortar table: rows (coarse element, direction,
    fine element, fine direction | perm << 3, size_a, size_b) with the mortar sizes
    in the coarse element's frame and perm the face permutation that takes a mortar
    point (a, b) of the coarse frame to the fine element's face point (0 between
    aligned blocks; bits as in connectivity_from_geometry)."""
    from scipy.spatial import cKDTree
    free = [(e, d) for e in range(nbr.shape[0]) for d in range(6) if nbr[e, d] == -1]
    if not free:
        return np.zeros((0, 6), dtype=np.int32)

    def face_point(d, a, b):
        xi = np.zeros(3)
        xi[d // 2] = 1.0 if d % 2 else -1.0
        fd = [x for x in range(3) if x != d // 2]
        xi[fd[0]], xi[fd[1]] = a, b
        return xi

    def phys(e, d, a, b):
        return dom.map_points(e, face_point(d, a, b)[:, None])[0][:, 0]

    def close(p, q):
        return np.linalg.norm(p - q) <= tol * max(np.linalg.norm(q), 1e-300)
    centers = np.array([phys(e, d, 0.0, 0.0) for e, d in free])
    tree = cKDTree(centers)
    # sub-face = (size code, centre, half width) per face dimension
    pieces = {0: [(MORTAR_FULL, 0.0, 1.0)],
              1: [(MORTAR_LOWER_HALF, -0.5, 0.5), (MORTAR_UPPER_HALF, 0.5, 0.5)]}
    mortars = []
    for k, (e, d) in enumerate(free):
        for split in ((1, 1), (1, 0), (0, 1)):
            found = []
            for sa, ca, ha in pieces[split[0]]:
                for sb, cb, hb in pieces[split[1]]:
                    q = phys(e, d, ca, cb)
                    dist, j = tree.query(q)
                    if dist < tol * max(np.linalg.norm(q), 1e-300) and j != k:
                        found.append((sa, ca, ha, sb, cb, hb, free[j]))
            if len(found) != len(pieces[split[0]]) * len(pieces[split[1]]):
                continue
            for sa, ca, ha, sb, cb, hb, (e2, d2) in found:
                # three corners of the mortar, in the coarse frame, tell the eight face
                # permutations apart
                code = None
                for perm in range(8):
                    ok = True
                    for a, b in ((-1.0, -1.0), (1.0, -1.0), (-1.0, 1.0)):
                        fa, fb = (b, a) if perm & 1 else (a, b)
                        if perm & 2:
                            fa = -fa
                        if perm & 4:
                            fb = -fb
                        if not close(phys(e2, d2, fa, fb), phys(e, d, ca + ha * a, cb + hb * b)):
                            ok = False
                            break
                    if ok:
                        code = perm
                        break
                if code is None:
                    raise ValueError(f"faces ({e},{d}) and ({e2},{d2}) share a centre but not "
                                     "their corners")
                mortars.append((e, d, e2, d2 | (code << 3), sa, sb))
            break
    for e, d, e2, d2, _, _ in mortars:
        nbr[e, d] = HANGING
        nbr[e2, d2 & 7] = HANGING
    return np.asarray(mortars, dtype=np.int32).reshape(-1, 6)


class Partition:
    """Contiguous split of the (Z-curve ordered) element list over `world`
    ranks and the halo bookkeeping of one rank.

    Local element order: interior elements (all neighbours local) first, then
    boundary elements, so that the interior range can run while the halo is in
    flight.  Ghost faces are numbered per peer in the order of the *receiver's*
    (local element, direction) list; the sender packs in the same order.
    """

    def __init__(self, neighbors: np.ndarray, world: int, rank: int,
                 boundary_slots=False, neighbor_direction=None, face_permutation=None,
                 mortars=None):
        """boundary_slots: give external faces (neighbour -1) of local elements a
        ghost slot after the exchanged ones, to be filled with the exterior state
        of a ghost boundary condition (DirichletAnalytic); True = every external
        face, or a predicate (global element, direction) -> bool.
        neighbor_direction / face_permutation [n_elements, 6]: orientation of
        non-aligned neighbours (default: aligned, neighbour direction d ^ 1).
        mortars [n, 6]: global table of the non-conforming mortars (coarse element,
        direction, fine element, direction, size_a, size_b).  local_mortars is the
        same table in local numbering, rows in the global order; a side that lives
        on another rank is a ghost slot, written -(slot + 2), which receives the
        remote face like any cut face (one slot per mortar)."""
        ne = neighbors.shape[0]
        mortars = np.zeros((0, 6), dtype=np.int64) if mortars is None else np.asarray(mortars)
        if neighbor_direction is None:
            neighbor_direction = np.tile((np.arange(6) ^ 1).astype(np.int32), (ne, 1))
            face_permutation = np.zeros((ne, 6), dtype=np.int32)
        self.oriented = bool((neighbor_direction != (np.arange(6) ^ 1)[None, :]).any()
                             or face_permutation.any())
        bounds = [(ne * r) // world for r in range(world + 1)]
        owner = np.zeros(ne, dtype=np.int64)
        for r in range(world):
            owner[bounds[r]:bounds[r + 1]] = r
        mine = np.arange(bounds[rank], bounds[rank + 1])
        if len(mortars) == 0 and (neighbors == HANGING).any():
            raise ValueError("hanging faces without a mortar table")
        nb_mine = neighbors[mine]
        remote = (nb_mine >= 0) & (owner[np.clip(nb_mine, 0, ne - 1)] != rank)
        is_boundary = remote.any(axis=1)
        # a mortar group (all mortars of one coarse face) with a member on another
        # rank needs halo data: every local participant is a boundary element
        group_remote = {}
        for ec, dc, ef, df, _, _ in mortars.tolist():
            key = (ec, dc)
            group_remote[key] = group_remote.get(key, False) or owner[ec] != owner[ef]
        for ec, dc, ef, df, _, _ in mortars.tolist():
            if group_remote[(ec, dc)]:
                for g in (ec, ef):
                    if owner[g] == rank:
                        is_boundary[g - bounds[rank]] = True
        order = np.concatenate([mine[~is_boundary], mine[is_boundary]])
        self.world, self.rank = world, rank
        self.global_ids = order                      # local -> global
        self.n_local = len(order)
        self.n_interior = int((~is_boundary).sum())
        g2l = {int(g): i for i, g in enumerate(order)}
        # receive list: (peer, neighbour global element, neighbour direction, local e, d)
        recv = []
        local_nb = np.full((self.n_local, 6), -1, dtype=np.int32)
        self.local_neighbor_direction = np.tile((np.arange(6) ^ 1).astype(np.int32),
                                                (self.n_local, 1))
        self.local_face_permutation = np.zeros((self.n_local, 6), dtype=np.int32)
        for le, g in enumerate(order):
            for d in range(6):
                v = int(neighbors[g, d])
                if v == HANGING:
                    local_nb[le, d] = HANGING
                if v < 0:
                    continue
                self.local_neighbor_direction[le, d] = neighbor_direction[g, d]
                self.local_face_permutation[le, d] = face_permutation[g, d]
                if owner[v] == rank:
                    local_nb[le, d] = g2l[v]
                else:
                    recv.append((int(owner[v]), v, int(neighbor_direction[g, d]), int(g), d,
                                 None))
        # faces of remote mortar partners: one ghost slot per mortar
        # (fine direction may carry the face permutation of non-aligned blocks in its
        # upper bits: every face travels in its owner's frame, so only the row keeps it)
        for m, (ec, dc, ef, df, _, _) in enumerate(mortars.tolist()):
            df &= 7
            if owner[ec] == rank and owner[ef] != rank:
                recv.append((int(owner[ef]), ef, df, ec, dc, m))
            elif owner[ef] == rank and owner[ec] != rank:
                recv.append((int(owner[ec]), ec, dc, ef, df, m))
        # canonical order shared by both sides: by (peer, global element of the
        # SENDER, sender direction, global element of the receiver, its direction)
        recv.sort(key=lambda t: t[:5])
        self.recv_counts = [0] * world
        mortar_slot = {}
        for slot, (peer, v, dn, g, d, m) in enumerate(recv):
            if m is None:
                local_nb[g2l[g], d] = -(slot + 2)
            else:
                mortar_slot[m] = slot
            self.recv_counts[peer] += 1
        self.n_recv = len(recv)
        rows = []
        for m, (ec, dc, ef, df, sa, sb) in enumerate(mortars.tolist()):
            if owner[ec] != rank and owner[ef] != rank:
                continue
            lc = g2l[ec] if owner[ec] == rank else -(mortar_slot[m] + 2)
            lf = g2l[ef] if owner[ef] == rank else -(mortar_slot[m] + 2)
            rows.append((lc, dc, lf, df, sa, sb))
        self.local_mortars = np.asarray(rows, dtype=np.int32).reshape(-1, 6)
        self.external_faces = []  # (local element, direction, slot)
        if boundary_slots:
            wanted = boundary_slots if callable(boundary_slots) else (lambda g, d: True)
            for le, g in enumerate(order):
                for d in range(6):
                    if int(neighbors[g, d]) == -1 and wanted(int(g), d):
                        slot = self.n_recv + len(self.external_faces)
                        self.external_faces.append((le, d, slot))
                        local_nb[le, d] = -(slot + 2)
        self.n_ghost = self.n_recv + len(self.external_faces)
        self.local_neighbors = local_nb
        # send list: my faces that some peer needs = faces of my elements whose
        # neighbour is remote; ordered by (peer, my global element, my direction)
        send = []
        for le, g in enumerate(order):
            for d in range(6):
                v = int(neighbors[g, d])
                if v >= 0 and owner[v] != rank:
                    # the receiver is the neighbour, seeing us through its direction
                    send.append((int(owner[v]), int(g), d, v, int(neighbor_direction[g, d]), le))
        for ec, dc, ef, df, _, _ in mortars.tolist():
            df &= 7
            if owner[ec] == rank and owner[ef] != rank:
                send.append((int(owner[ef]), ec, dc, ef, df, g2l[ec]))
            elif owner[ef] == rank and owner[ec] != rank:
                send.append((int(owner[ec]), ef, df, ec, dc, g2l[ef]))
        send.sort(key=lambda t: t[:5])
        self.send_counts = [0] * world
        self.send_map = np.zeros((len(send), 2), dtype=np.int32)
        for slot, (peer, g, d, gr, dr, le) in enumerate(send):
            self.send_map[slot] = (le, d)
            self.send_counts[peer] += 1
        assert len(send) == len(recv) or world > 1  # symmetric on periodic bricks

    def reorder(self, perm):
        """Single-rank partition: put the local elements into the order perm (new -> old), e.g.
        sorted by step-size level for local time stepping; every table follows."""
        if self.world != 1:
            raise ValueError("Partition.reorder: single-rank partitions only")
        perm = np.asarray(perm)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        self.global_ids = self.global_ids[perm]
        nb = self.local_neighbors[perm].copy()
        m = nb >= 0
        nb[m] = inv[nb[m]]
        self.local_neighbors = nb.astype(np.int32)
        self.local_neighbor_direction = self.local_neighbor_direction[perm]
        self.local_face_permutation = self.local_face_permutation[perm]
        if len(self.local_mortars):
            lm = self.local_mortars.copy()
            lm[:, 0], lm[:, 2] = inv[lm[:, 0]], inv[lm[:, 2]]
            self.local_mortars = lm.astype(np.int32)
        self.external_faces = [(int(inv[le]), d, slot) for le, d, slot in self.external_faces]
        self.send_map = np.zeros((0, 2), dtype=np.int32)
