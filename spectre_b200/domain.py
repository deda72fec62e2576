"""Host-side domain data model for the DG path: a periodic/non-periodic Brick
of 2^L elements per dimension, elements ordered along the per-block Z-curve
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


class Partition:
    """Contiguous split of the (Z-curve ordered) element list over `world`
    ranks and the halo bookkeeping of one rank.

    Local element order: interior elements (all neighbours local) first, then
    boundary elements, so that the interior range can run while the halo is in
    flight.  Ghost faces are numbered per peer in the order of the *receiver's*
    (local element, direction) list; the sender packs in the same order.
    """

    def __init__(self, neighbors: np.ndarray, world: int, rank: int,
                 boundary_slots: bool = False):
        """boundary_slots: give every external face (neighbour -1) of a local
        element a ghost slot after the exchanged ones, to be filled with the
        exterior state of a ghost boundary condition (DirichletAnalytic)."""
        ne = neighbors.shape[0]
        bounds = [(ne * r) // world for r in range(world + 1)]
        owner = np.zeros(ne, dtype=np.int64)
        for r in range(world):
            owner[bounds[r]:bounds[r + 1]] = r
        mine = np.arange(bounds[rank], bounds[rank + 1])
        nb_mine = neighbors[mine]
        remote = (nb_mine >= 0) & (owner[np.clip(nb_mine, 0, ne - 1)] != rank)
        is_boundary = remote.any(axis=1)
        order = np.concatenate([mine[~is_boundary], mine[is_boundary]])
        self.world, self.rank = world, rank
        self.global_ids = order                      # local -> global
        self.n_local = len(order)
        self.n_interior = int((~is_boundary).sum())
        g2l = {int(g): i for i, g in enumerate(order)}
        # receive list: (peer, neighbour global element, neighbour direction, local e, d)
        recv = []
        local_nb = np.full((self.n_local, 6), -1, dtype=np.int32)
        for le, g in enumerate(order):
            for d in range(6):
                v = int(neighbors[g, d])
                if v < 0:
                    continue
                if owner[v] == rank:
                    local_nb[le, d] = g2l[v]
                else:
                    recv.append((int(owner[v]), v, d ^ 1, le, d))
        # canonical order shared by both sides: by (peer, global element of the
        # SENDER, sender direction)
        recv.sort(key=lambda t: (t[0], t[1], t[2]))
        self.recv_counts = [0] * world
        for slot, (peer, v, dn, le, d) in enumerate(recv):
            local_nb[le, d] = -(slot + 2)
            self.recv_counts[peer] += 1
        self.n_recv = len(recv)
        self.external_faces = []  # (local element, direction, slot)
        if boundary_slots:
            for le, g in enumerate(order):
                for d in range(6):
                    if int(neighbors[g, d]) == -1:
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
                    send.append((int(owner[v]), int(g), d, le))
        send.sort(key=lambda t: (t[0], t[1], t[2]))
        self.send_counts = [0] * world
        self.send_map = np.zeros((len(send), 2), dtype=np.int32)
        for slot, (peer, g, d, le) in enumerate(send):
            self.send_map[slot] = (le, d)
            self.send_counts[peer] += 1
        assert len(send) == len(recv) or world > 1  # symmetric on periodic bricks
