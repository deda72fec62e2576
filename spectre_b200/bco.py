"""Multi-block domains from coordinate maps, and the BinaryCompactObject creator.

Host-side set-up only: the hot path consumes inertial coordinates,
inverse Jacobians and the neighbour / orientation / mortar tables; this module produces them
for domains given as a list of block maps.  The maps are restated from the reference's
forward maps (`Domain/CoordinateMaps/Wedge.cpp:250-537`, `Frustum.cpp:30-213`,
`Interval.cpp:66-78`); their Jacobians are taken by complex-step differentiation of the
forward map (exact to rounding for these analytic maps), not from the reference's
hand-written Jacobians.  Block layout of `BinaryCompactObject`:
`Domain/Creators/BinaryCompactObject.cpp:84-140,420-552` with
`sph_wedge_coordinate_maps` / `frustum_coordinate_maps` of `DomainHelpers.cpp:595-840`.
"""
from __future__ import annotations

import numpy as np

from . import lib
from .domain import (WEDGE_ORIENTATIONS, connectivity_from_geometry, element_id,
                     find_hanging_faces, z_curve_index)


def _rotate(orientation, src):
    """discrete_rotation(orientation, src): x[i] = sign_i * src[dim_i]"""
    return [orientation[i][1] * src[orientation[i][0]] for i in range(3)]


def _rotate_inverse(orientation, x):
    src = [0.0, 0.0, 0.0]
    for i in range(3):
        src[orientation[i][0]] = orientation[i][1] * x[i]
    return src


class Wedge3:
    """CoordinateMaps::Wedge<3> without a focal offset (Wedge.cpp:67-130, 250-537)."""

    def __init__(self, radius_inner, radius_outer, sphericity_inner, sphericity_outer,
                 orientation, equiangular=True, halves="Both", distribution="Linear",
                 opening_angles=(0.5 * np.pi, 0.5 * np.pi), adapted_equiangular=True):
        self.o = WEDGE_ORIENTATIONS[orientation] if isinstance(orientation, int) else orientation
        self.equiangular, self.halves, self.distribution = equiangular, halves, distribution
        self.opening = tuple(opening_angles)
        self.opening_dist = self.opening if adapted_equiangular else (0.5 * np.pi, 0.5 * np.pi)
        ri, ro, si, so = radius_inner, radius_outer, sphericity_inner, sphericity_outer
        if distribution == "Linear":
            s3 = np.sqrt(3.0)
            self.sphere_zero = 0.5 * (so * ro + si * ri)
            self.sphere_rate = 0.5 * (so * ro - si * ri)
            self.frustum_zero = 0.5 / s3 * ((1.0 - so) * ro + (1.0 - si) * ri)
            self.frustum_rate = 0.5 / s3 * ((1.0 - so) * ro - (1.0 - si) * ri)
        elif distribution == "Logarithmic":
            if si != 1.0 or so != 1.0:
                raise ValueError("Logarithmic radial distribution needs spherical surfaces")
            self.sphere_zero = 0.5 * np.log(ro * ri)
            self.sphere_rate = 0.5 * np.log(ro / ri)
            self.frustum_zero = self.frustum_rate = 0.0
        else:
            raise ValueError(f"unsupported radial distribution {distribution}")

    def _cap(self, index, v):
        if not self.equiangular:
            return v
        return (np.tan(0.5 * self.opening[index]) * np.tan(0.5 * self.opening_dist[index] * v)
                / np.tan(0.5 * self.opening_dist[index]))

    def __call__(self, xi, eta, zeta):
        if self.halves == "UpperOnly":
            xi = 0.5 * (xi + 1.0)
        elif self.halves == "LowerOnly":
            xi = 0.5 * (xi - 1.0)
        cap0, cap1 = self._cap(0, xi), self._cap(1, eta)
        one_over_rho = 1.0 / np.sqrt(1.0 + cap0 * cap0 + cap1 * cap1)
        if self.distribution == "Linear":
            z = ((self.sphere_zero + self.sphere_rate * zeta) * one_over_rho
                 + (self.frustum_zero + self.frustum_rate * zeta))
        else:
            z = np.exp(self.sphere_zero + self.sphere_rate * zeta) * one_over_rho
        return _rotate(self.o, [z * cap0, z * cap1, z])


class Frustum:
    """CoordinateMaps::Frustum (Frustum.cpp:30-213)."""

    def __init__(self, face_vertices, lower_bound, upper_bound, orientation, equiangular=True,
                 distribution="Linear", distribution_value=None, sphericity=0.0,
                 transition_phi=0.0, opening_angle=0.5 * np.pi):
        self.o = WEDGE_ORIENTATIONS[orientation] if isinstance(orientation, int) else orientation
        (lxl, lyl), (uxl, uyl), (lxu, lyu), (uxu, uyu) = face_vertices
        self.equiangular, self.distribution, self.sphericity = equiangular, distribution, sphericity
        self.sigma_x = 0.25 * (lxu + uxu + lxl + uxl)
        self.delta_x_zeta = 0.25 * (lxu + uxu - lxl - uxl)
        self.delta_x_xi = 0.25 * (uxu - lxu + uxl - lxl)
        self.delta_x_xi_zeta = 0.25 * (uxu - lxu - uxl + lxl)
        self.sigma_y = 0.25 * (lyu + uyu + lyl + uyl)
        self.delta_y_zeta = 0.25 * (lyu + uyu - lyl - uyl)
        self.delta_y_eta = 0.25 * (uyu - lyu + uyl - lyl)
        self.delta_y_eta_zeta = 0.25 * (uyu - lyu - uyl + lyl)
        self.sigma_z = 0.5 * (upper_bound + lower_bound)
        self.delta_z_zeta = 0.5 * (upper_bound - lower_bound)
        self.phi = transition_phi
        self.half_opening = 0.5 * opening_angle
        self.radius = np.sqrt(max(abs(uxu), abs(uxl), abs(lxu), abs(lxl)) ** 2
                              + max(abs(uyu), abs(uyl), abs(lyu), abs(lyl)) ** 2
                              + max(abs(upper_bound), abs(lower_bound)) ** 2)
        inner_radius = np.sqrt(max(abs(uxl), abs(lxl)) ** 2 + max(abs(uyl), abs(lyl)) ** 2
                               + lower_bound ** 2)
        if distribution == "Projective":
            w_delta = distribution_value if distribution_value is not None else np.sqrt(
                ((uxl - lxl) * (uyl - lyl)) / ((uxu - lxu) * (uyu - lyu)))
            self.w_plus, self.w_minus = w_delta + 1.0, w_delta - 1.0
        elif distribution == "Logarithmic":
            self.singularity = (distribution_value if distribution_value is not None else
                                -(self.radius + inner_radius) / (self.radius - inner_radius))
        elif distribution != "Linear":
            raise ValueError(distribution)

    def __call__(self, xi, eta, zeta):
        if self.distribution == "Projective":
            cap_zeta = (self.w_minus + self.w_plus * zeta) / (self.w_plus + self.w_minus * zeta)
        elif self.distribution == "Linear":
            cap_zeta = zeta
        else:
            # Interval(-1, 1, -1, 1, Logarithmic, singularity) (Interval.cpp:66-78)
            s = self.singularity
            zero = 0.5 * np.log((1.0 - s) * (-1.0 - s))
            rate = 0.5 * np.log((1.0 - s) / (-1.0 - s))
            sign = 1.0 if -1.0 > s else -1.0
            cap_zeta = sign * np.exp(zero + rate * zeta) + s
        if self.equiangular:
            cap_xi_zero = np.tan(0.25 * np.pi * xi)
            opp = 1.0 + self.phi * self.phi
            cap_xi_upper = (opp / np.tan(self.half_opening)
                            * np.tan(self.half_opening * (xi + self.phi) / opp) - self.phi)
            cap_eta = np.tan(0.25 * np.pi * eta)
        else:
            cap_xi_zero = cap_xi_upper = xi
            cap_eta = eta
        cap_xi = 0.5 * (1.0 + cap_zeta) * cap_xi_upper + 0.5 * (1.0 - cap_zeta) * cap_xi_zero
        x = (self.sigma_x + self.delta_x_xi * cap_xi
             + (self.delta_x_zeta + self.delta_x_xi_zeta * cap_xi) * cap_zeta)
        y = (self.sigma_y + self.delta_y_eta * cap_eta
             + (self.delta_y_zeta + self.delta_y_eta_zeta * cap_eta) * cap_zeta)
        z = self.sigma_z + self.delta_z_zeta * cap_zeta
        if self.sphericity > 0.0:
            ux = (self.sigma_x + self.delta_x_xi * cap_xi_upper
                  + (self.delta_x_zeta + self.delta_x_xi_zeta * cap_xi_upper))
            uy = (self.sigma_y + self.delta_y_eta * cap_eta
                  + (self.delta_y_zeta + self.delta_y_eta_zeta * cap_eta))
            uz = self.sigma_z + self.delta_z_zeta
            ur = np.sqrt(ux * ux + uy * uy + uz * uz)
            c = 0.5 * self.sphericity * (1.0 + cap_zeta) * (self.radius / ur - 1.0)
            x, y, z = x + c * ux, y + c * uy, z + c * uz
        return _rotate(self.o, [x, y, z])


class Translated:
    def __init__(self, inner, shift):
        self.inner, self.shift = inner, shift

    def __call__(self, xi, eta, zeta):
        x = self.inner(xi, eta, zeta)
        return [x[i] + self.shift[i] for i in range(3)]


class MultiBlockDomain:
    """Blocks given as maps (xi, eta, zeta) -> [x, y, z] (numpy, complex-safe), each refined
    to 2^l elements per dimension; block-major element order, Z-curve inside a block
    (ElementDistribution.hpp:33-47).  Same interface as domain.SphericalShell."""

    def __init__(self, block_maps, block_levels, N, block_names=None):
        self.N, self.n = int(N), int(N) ** 3
        self.block_maps = list(block_maps)
        self.block_levels = [tuple(int(v) for v in lev) for lev in block_levels]
        self.block_names = list(block_names) if block_names else [""] * len(self.block_maps)
        self.n_blocks = len(self.block_maps)
        self.cells = []
        for b, lev in enumerate(self.block_levels):
            nx, ny, nz = (2 ** v for v in lev)
            cells = [(ix, iy, iz) for iz in range(nz) for iy in range(ny) for ix in range(nx)]
            cells.sort(key=lambda c: z_curve_index(c[0], c[1], c[2], lev))
            self.cells += [(b, c) for c in cells]
        self.n_elements = len(self.cells)
        self.xi, self.weights = lib.collocation_points_and_weights(self.N)
        self._conn = None

    def element_ids(self):
        return [element_id(b, c, self.block_levels[b]) for b, c in self.cells]

    def _with_order(self, N):
        return MultiBlockDomain(self.block_maps, self.block_levels, N, self.block_names)

    def map_points(self, e, xi):
        """element-logical xi [3, m] -> (x [3, m], jacobian [3, 3, m] w.r.t. the element-logical
        coordinates), the Jacobian by complex-step differentiation"""
        b, cell = self.cells[e]
        blk, half = [], []
        for d in range(3):
            h = 2.0 / 2 ** self.block_levels[b][d]
            blk.append(-1.0 + h * cell[d] + 0.5 * h * (np.asarray(xi[d], float) + 1.0))
            half.append(0.5 * h)
        fmap = self.block_maps[b]
        x = np.array([np.real(v) * np.ones_like(blk[0]) for v in fmap(*blk)])
        jac = np.empty((3, 3) + blk[0].shape)
        step = 1e-30
        for j in range(3):
            arg = [v.astype(complex) for v in blk]
            arg[j] = arg[j] + 1j * step
            out = fmap(*arg)
            for i in range(3):
                jac[i, j] = np.imag(out[i] * np.ones_like(arg[0])) / step * half[j]
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
            inv = np.linalg.inv(np.moveaxis(jac, -1, 0))
            for jh in range(3):
                for i in range(3):
                    Jinv[k, jh + 3 * i] = inv[:, jh, i]
        return X, Jinv

    def coords(self, ids=None):
        return self._geometry(list(range(self.n_elements) if ids is None else ids))[0]

    def inverse_jacobian(self, ids=None):
        return self._geometry(list(range(self.n_elements) if ids is None else ids))[1]

    def _connectivity(self):
        if self._conn is None:
            corners = self._with_order(2)
            nbr, nd, perm = connectivity_from_geometry(corners.coords(), 2)
            mortars = find_hanging_faces(corners, nbr)
            self._conn = (nbr, nd, perm, mortars)
        return self._conn

    def neighbors(self):
        return self._connectivity()[0]

    def neighbor_orientations(self):
        c = self._connectivity()
        return c[1], c[2]

    def mortars(self):
        return self._connectivity()[3]

    def block_of(self, e):
        return self.cells[e][0]


class BinaryCompactObject(MultiBlockDomain):
    """domain::creators::BinaryCompactObject with both objects excised, CubeScale 1 and no
    centre-of-mass offset (BinaryCompactObject.cpp:84-140, 420-552): per object six spherical
    wedges (inner radius -> outer radius) and six wedges from that sphere to the object's
    cube, ten bulged frustums from the two abutting cubes to the envelope sphere, ten
    (half-)wedges from the envelope to the outer sphere: 44 blocks.  refinement: one level
    (all blocks, all dimensions) or a dict block group -> (l_xi, l_eta, l_zeta) with the
    groups ObjectAShell, ObjectACube, ObjectBShell, ObjectBCube, Envelope, OuterShell
    (Inspiral.yaml:95-101)."""

    GROUPS = ("ObjectAShell", "ObjectACube", "ObjectBShell", "ObjectBCube", "Envelope",
              "OuterShell")

    def __init__(self, x_coord_a, x_coord_b, inner_radius_a, outer_radius_a, inner_radius_b,
                 outer_radius_b, envelope_radius, outer_radius, refinement, N,
                 opening_angle_degrees=90.0, equiangular=True, object_logarithmic=True,
                 envelope_distribution="Logarithmic", outer_shell_distribution="Linear"):
        if x_coord_a <= 0.0 or x_coord_b >= 0.0:
            raise ValueError("ObjectA sits at positive x, ObjectB at negative x")
        opening = np.pi * opening_angle_degrees / 180.0
        length_inner = x_coord_a - x_coord_b                  # CubeScale 1
        length_outer = 2.0 * envelope_radius / np.sqrt(2.0 + np.tan(0.5 * opening) ** 2)
        translation = 0.5 * (x_coord_a + x_coord_b)
        if envelope_radius <= length_inner * np.sqrt(3.0):
            raise ValueError("the envelope radius is too small: the frustums would be malformed")
        if envelope_radius >= outer_radius:
            raise ValueError("the outer radius must be larger than the envelope radius")
        maps, names = [], []
        for tag, xc, r_in, r_out in (("A", x_coord_a, inner_radius_a, outer_radius_a),
                                     ("B", x_coord_b, inner_radius_b, outer_radius_b)):
            if not r_in < r_out < 0.5 * length_inner:
                raise ValueError(f"Object{tag}: need inner radius < outer radius < half the cube")
            shift = (xc, 0.0, 0.0)
            dist = "Logarithmic" if object_logarithmic else "Linear"
            for w in range(6):
                maps.append(Translated(Wedge3(r_in, r_out, 1.0, 1.0, w, equiangular, "Both", dist),
                                       shift))
                names.append(f"Object{tag}Shell")
            for w in range(6):
                maps.append(Translated(Wedge3(r_out, np.sqrt(3.0) * 0.5 * length_inner, 1.0, 0.0,
                                              w, equiangular), shift))
                names.append(f"Object{tag}Cube")
        # ten frustums (DomainHelpers.cpp:734-840)
        lower, top, stretch = 0.5 * length_inner, 0.5 * length_outer, np.tan(0.5 * opening)
        origin_preimage = [-translation, 0.0, 0.0]
        value = (length_inner / length_outer if envelope_distribution == "Projective" else
                 -(length_outer + length_inner) / (length_outer - length_inner))
        for i in range(4):
            disp = _rotate_inverse(WEDGE_ORIENTATIONS[i], origin_preimage)
            maps.append(Frustum([(-2.0 * lower - disp[0], -lower - disp[1]),
                                 (-disp[0], lower - disp[1]), (stretch * -top, -top), (0.0, top)],
                                lower - disp[2], top, i, equiangular, envelope_distribution, value,
                                1.0, -1.0, opening))
            maps.append(Frustum([(-disp[0], -lower - disp[1]),
                                 (2.0 * lower - disp[0], lower - disp[1]), (0.0, -top),
                                 (stretch * top, top)],
                                lower - disp[2], top, i, equiangular, envelope_distribution, value,
                                1.0, 1.0, opening))
        for i in (4, 5):
            disp = _rotate_inverse(WEDGE_ORIENTATIONS[i], origin_preimage)
            maps.append(Frustum([(-lower - disp[0], -lower - disp[1]),
                                 (lower - disp[0], lower - disp[1]), (-top, -top), (top, top)],
                                2.0 * lower - disp[2], stretch * top, i, equiangular,
                                envelope_distribution, value, 1.0, 0.0, 0.5 * np.pi))
        names += ["Envelope"] * 10
        # outer shell: half wedges around the x axis, full end caps (DomainHelpers.cpp:692-720)
        for i in range(4):
            for halves in ("LowerOnly", "UpperOnly"):
                maps.append(Wedge3(envelope_radius, outer_radius, 1.0, 1.0, i, equiangular, halves,
                                   outer_shell_distribution, (opening, 0.5 * np.pi)))
        cap = np.pi - opening
        for i in (4, 5):
            maps.append(Wedge3(envelope_radius, outer_radius, 1.0, 1.0, i, equiangular, "Both",
                               outer_shell_distribution, (cap, cap), adapted_equiangular=False))
        names += ["OuterShell"] * 10
        if isinstance(refinement, int):
            refinement = {g: (refinement,) * 3 for g in self.GROUPS}
        levels = [tuple(refinement[g]) for g in names]
        self.parameters = dict(x_coord_a=x_coord_a, x_coord_b=x_coord_b,
                               inner_radius_a=inner_radius_a, inner_radius_b=inner_radius_b,
                               envelope_radius=envelope_radius, outer_radius=outer_radius)
        super().__init__(maps, levels, N, names)

    def _with_order(self, N):
        clone = MultiBlockDomain(self.block_maps, self.block_levels, N, self.block_names)
        return clone

    def external_boundary(self, e, d):
        """'excision_a', 'excision_b' or 'outer' for an external face"""
        name = self.block_names[self.cells[e][0]]
        if name == "ObjectAShell" and d == 4:
            return "excision_a"
        if name == "ObjectBShell" and d == 4:
            return "excision_b"
        if name == "OuterShell" and d == 5:
            return "outer"
        raise ValueError(f"face {d} of a {name} block is not an external boundary")
