"""Domains whose blocks have different numbers of grid points (p-refinement): one context per
N ("element classes"), the faces between classes are p-mortars.  Host-side
driver over the C-ABI: builds the per-class neighbour tables, the p-mortar tables and the halo
maps from the domain's global connectivity and steps all contexts together, the loop of
INTEGRATION.md "Elements with different numbers of grid points"."""
from __future__ import annotations

import numpy as np

from . import lib


class PRefinedEvolution:
    """dom: a multi-block domain (bco.MultiBlockDomain interface: cells, block_of,
    _with_order, neighbors, neighbor_orientations; conforming in h).  points_of_block: N per
    block.  External faces carry DirichletAnalytic with `initial_data` (static data).  GH:
    AnalyticChristoffel gauge of the initial data."""

    def __init__(self, system, dom, points_of_block, initial_data, static_values,
                 stepper=lib.STEPPER_ADAMS_BASHFORTH, order=3, dt=1e-3, t0=0.0, device=0):
        self.system, self.dom = system, dom
        if len(dom.mortars()):
            raise ValueError("p-refinement needs a domain that conforms in h")
        nbr, (nd, perm) = dom.neighbors(), dom.neighbor_orientations()
        n_of = np.array([points_of_block[dom.block_of(e)] for e in range(dom.n_elements)])
        self.Ns = sorted(set(int(v) for v in n_of))
        self.ids = [np.nonzero(n_of == N)[0] for N in self.Ns]          # global ids per class
        cls_of = {int(N): k for k, N in enumerate(self.Ns)}
        local = np.empty(dom.n_elements, dtype=np.int64)
        for ids in self.ids:
            local[ids] = np.arange(len(ids))
        C = 50 if system == lib.SYSTEM_GH else 5
        self.C = C
        # per class: neighbour / orientation tables, p-mortar table, halo map, external faces
        self.tables = []
        for k, ids in enumerate(self.ids):
            ln = np.full((len(ids), 6), -1, dtype=np.int32)
            ldir = np.tile((np.arange(6) ^ 1).astype(np.int32), (len(ids), 1))
            lperm = np.zeros((len(ids), 6), dtype=np.int32)
            pm, send, ext = [], [], []
            for le, g in enumerate(ids):
                for d in range(6):
                    g2 = nbr[g, d]
                    if g2 < 0:
                        ln[le, d] = -(len(ext) + 2)          # DirichletAnalytic ghost slot
                        ext.append((le, d))
                    elif n_of[g2] == n_of[g]:
                        ln[le, d], ldir[le, d], lperm[le, d] = local[g2], nd[g, d], perm[g, d]
                    else:
                        ln[le, d] = lib.P_MORTAR
                        pm.append((le, d, int(n_of[g2]), int(nd[g, d]) | (int(perm[g, d]) << 3),
                                   int(g2)))
                        send.append((le, d))
            self.tables.append(dict(nbr=ln, nbr_dir=ldir, face_perm=lperm, pm=pm, send=send,
                                    ext=ext))
        # link the two sides: face i of class a receives halo slot j of class b
        slot_of = [{(int(ids[le]), d): j for j, (le, d) in enumerate(t["send"])}
                   for ids, t in zip(self.ids, self.tables)]
        self.transfers = {}      # (src class, dst class) -> (src slots, dst faces)
        for a, t in enumerate(self.tables):
            for i, (le, d, nb_points, code, g2) in enumerate(t["pm"]):
                b = cls_of[nb_points]
                src, dst = self.transfers.setdefault((b, a), ([], []))
                src.append(slot_of[b][(g2, code & 7)])
                dst.append(i)
        self.ctxs, self.x, self.J, self.stat, self.u0 = [], [], [], [], []
        for k, (N, ids, t) in enumerate(zip(self.Ns, self.ids, self.tables)):
            geo = dom._with_order(N)
            x, J = geo.coords(ids), geo.inverse_jacobian(ids)
            stat = np.empty((len(ids), len(static_values), N ** 3))
            for i, v in enumerate(static_values):
                stat[:, i] = v(x) if callable(v) else v
            u0 = initial_data(x, t0)
            n_ghost = max(len(t["ext"]), len(t["send"]), 1)
            ctx = lib.Context(system, N, len(ids), n_ghost, device)
            ctx.set_geometry(J, x, t["nbr"])
            ctx.set_neighbor_orientations(t["nbr_dir"], t["face_perm"])
            ctx.set_static_fields(stat)
            if system == lib.SYSTEM_GH:
                ctx.set_gauge_analytic_christoffel(u0)
            ctx.set_state(u0)
            ctx.set_p_mortars([row[:4] for row in t["pm"]])
            ctx.set_halo_map(np.array(t["send"], dtype=np.int32).reshape(-1, 2))
            ctx.set_interior_count(0)
            if t["ext"]:
                ctx.set_boundary_ghost_data(0, self.boundary_ghost_data(k, x, J, stat, u0))
            ctx.set_stepper(stepper, order, t0, dt)
            self.ctxs.append(ctx)
            self.x.append(x), self.J.append(J), self.stat.append(stat), self.u0.append(u0)

    def boundary_ghost_data(self, k, x, J, stat, u0):
        """[n_external][halo comps][N^2]: the analytic state on the face, then the interior
        element's inverse-Jacobian row and gammas (evolution.boundary_ghost_data)."""
        N, t = self.Ns[k], self.tables[k]
        f, C = N * N, self.C
        hc = C + 3 + (2 if self.system == lib.SYSTEM_GH else 1)
        q = np.arange(f)
        a, b = q % N, q // N
        out = np.zeros((len(t["ext"]), hc, f))
        for s, (le, d) in enumerate(t["ext"]):
            dim, fixed = d // 2, (N - 1 if d % 2 else 0)
            p = [fixed + N * (a + N * b), a + N * (fixed + N * b), a + N * (b + N * fixed)][dim]
            out[s, :C] = u0[le][:, p]
            for i in range(3):
                out[s, C + i] = J[le][dim + 3 * i, p]
            if self.system == lib.SYSTEM_GH:
                out[s, C + 3], out[s, C + 4] = stat[le][1, p], stat[le][2, p]
            else:
                out[s, C + 3] = stat[le][0, p]
        return out

    def compute_time_derivative(self, time):
        """one right-hand side of all classes (inside or outside a substep)"""
        for ctx in self.ctxs:
            ctx.pack_halo()
        for (src, dst), (slots, faces) in self.transfers.items():
            self.ctxs[dst].p_mortar_transfer_from(self.ctxs[src], slots, faces)
        for ctx in self.ctxs:
            ctx.compute_time_derivative_range(time, 0, ctx.n_elements)

    def take_steps(self, n):
        done = 0
        while done < n:
            times = [ctx.begin_substep() for ctx in self.ctxs]
            assert len(set(times)) == 1
            self.compute_time_derivative(times[0])
            done += [ctx.end_substep() for ctx in self.ctxs][0]

    def close(self):
        for ctx in self.ctxs:
            ctx.close()
