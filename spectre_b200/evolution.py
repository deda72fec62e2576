"""Evolution driver on top of the C-ABI: problem set-up for the BASELINE.json
configurations and the per-substep schedule, single- or multi-GPU.

Multi-GPU: the Z-curve ordered element list is cut into one
contiguous chunk per rank (spectre_b200.domain.Partition); per RHS the only
communication is the exchange of the cut mortar faces -- the counterpart of
the reference's send_data_for_fluxes / receive_boundary_data_global_time_stepping
(ComputeTimeDerivative.hpp:652-774, ApplyBoundaryCorrections.hpp:205-380),
except that raw face values (55 components) travel instead of the 134-component
packaged data.  Schedule per substep:

    pack_halo (ctx stream) -> NCCL send/recv (comm stream)
    faces + volume of interior elements (overlap the exchange)
    remaining faces (comm stream, after the halo) -> volume of boundary elements
    -> stepper update

The exchange and the whole schedule live inside libdgrhs.so (dgrhs_comm_init,
dgrhs_take_steps): torch.distributed only hands the 128-byte ncclUniqueId to the
other ranks.  The Python-driven schedule over torch.distributed (HaloExchange) is
kept for the gloo CPU tests, for time-dependent boundary data and as an A/B
reference (Evolution(..., native_exchange=False)).
"""
from __future__ import annotations

import numpy as np

from . import analytic, domain, lib


class _CudaArray:
    """Expose a raw device pointer through __cuda_array_interface__."""

    def __init__(self, ptr, nelem_f64):
        self.__cuda_array_interface__ = {
            "shape": (nelem_f64,), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
            "strides": None,
        }


class HaloExchange:
    """The one exchange step of the path: every cut mortar face travels to the
    rank that owns the neighbouring element (point-to-point, no collective).
    Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU
    tests); buffers are flat float64 tensors [n_ghost * per_face]."""

    def __init__(self, part, per_face, dist, group=None):
        self.dist, self.group = dist, group
        self.segments = []
        so = ro = 0
        for peer in range(part.world):
            ns, nr = part.send_counts[peer], part.recv_counts[peer]
            if ns or nr:
                self.segments.append((peer, so * per_face, ns * per_face, ro * per_face,
                                      nr * per_face))
            so += ns
            ro += nr

    def start(self, send, recv):
        dist = self.dist
        ops = []
        for peer, so, ns, ro, nr in self.segments:
            if nr:
                ops.append(dist.P2POp(dist.irecv, recv[ro:ro + nr], peer, group=self.group))
            if ns:
                ops.append(dist.P2POp(dist.isend, send[so:so + ns], peer, group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []


class Problem:
    """A BASELINE.json configuration; per-element data are built on demand for
    the element subset a rank owns (global arrays would not fit at 8 GPUs)."""

    def __init__(self, system, brick, initial_data, static_values, dirichlet_analytic=False,
                 analytic_christoffel_gauge=False, demand_outgoing=None, bjorhus=None):
        """brick: the domain (domain.Brick or domain.SphericalShell).
        static_values: one number or one callable(x) -> array per static field.
        dirichlet_analytic: external faces get the analytic solution as exterior
        state (DirichletAnalytic ghost boundary condition); True = all of them, or
        a predicate (global element, direction) -> bool.
        demand_outgoing: DemandOutgoingCharSpeeds on the remaining external faces
        (no correction, characteristic speeds checked); None = off.
        bjorhus: function (global element, direction) -> None / "ConstraintPreserving" /
        "ConstraintPreservingPhysical" selecting external faces with
        ConstraintPreservingBjorhus of that type.
        analytic_christoffel_gauge: AnalyticChristoffel gauge of the (static)
        analytic solution instead of the harmonic gauge."""
        self.bjorhus = bjorhus
        self.system, self.brick, self.N = system, brick, brick.N
        self._initial_data, self._static_values = initial_data, static_values
        self.dirichlet_analytic = dirichlet_analytic
        self.demand_outgoing = bool(demand_outgoing)
        self.analytic_christoffel_gauge = analytic_christoffel_gauge
        # non-aligned neighbours (multi-block domains)
        self.orientations = (brick.neighbor_orientations()
                             if hasattr(brick, "neighbor_orientations") else (None, None))
        # a time-dependent analytic solution on the boundary is re-evaluated at the
        # time of every RHS (DirichletAnalytic.cpp:80-96 passes `time`)
        self.boundary_time_dependent = False
        self.neighbors = brick.neighbors()
        # non-conforming (2:1) mortars of h-refined domains
        self.mortars = brick.mortars() if hasattr(brick, "mortars") else np.zeros((0, 6), int)

    def coords(self, ids=None):
        return self.brick.coords(ids)

    def inverse_jacobian(self, ids=None):
        return self.brick.inverse_jacobian(ids)

    def u0(self, ids=None, t=0.0):
        return self._initial_data(self.coords(ids), t)

    def static(self, ids=None):
        ne = self.brick.n_elements if ids is None else len(ids)
        out = np.empty((ne, len(self._static_values), self.brick.n))
        x = None
        for i, v in enumerate(self._static_values):
            if callable(v):
                x = self.coords(ids) if x is None else x
                out[:, i] = v(x)
            else:
                out[:, i] = v
        return out


def gh_gauge_wave_problem(refinement, N, lower=(0.0, 0.0, 0.0), upper=(1.0, 1.0, 1.0),
                          amplitude=0.1, wavelength=1.0, gammas=(1.0, -1.0, 1.0)):
    """BASELINE.json configs[1] (GaugeWave3D.yaml: Brick [0,1]^3 periodic,
    gamma0 = 1, gamma1 = -1, gamma2 = 1)."""
    brick = domain.Brick(lower, upper, refinement, N)
    return Problem(lib.SYSTEM_GH, brick,
                   lambda x, t: analytic.gauge_wave(x, t, amplitude, wavelength), gammas)


def gh_kerr_schild_problem(refinement, N, lower=(2.0, 2.0, 2.0), upper=(4.0, 4.0, 4.0),
                           mass=1.0):
    """Kerr-Schild black hole (M = 1, a = 0; KerrSchild.yaml:73-78) on a Brick
    that does not contain the singularity, DirichletAnalytic on all external
    faces, AnalyticChristoffel gauge, GaussianPlusConstant damping functions of
    KerrSchild.yaml:108-125 -- the Brick-lattice stand-in for BASELINE.json
    configs[2]/[3] (the Sphere domain with wedges is not built yet)."""
    brick = domain.Brick(lower, upper, refinement, N, periodic=(False, False, False))
    w = 11.313708499
    gam = (lambda x: analytic.gaussian_plus_constant(x, 0.001, 3.0, w),
           -1.0,
           lambda x: analytic.gaussian_plus_constant(x, 0.001, 1.0, w))
    return Problem(lib.SYSTEM_GH, brick, lambda x, t: analytic.kerr_schild(x, mass), gam,
                   dirichlet_analytic=True, analytic_christoffel_gauge=True)


def gh_kerr_schild_shell_problem(refinement, N, inner_radius=1.9, outer_radius=2.3,
                                 radial_partitioning=(), mass=1.0,
                                 inner_boundary="DirichletAnalytic",
                                 radial_distribution="Logarithmic", order="block",
                                 outer_boundary="DirichletAnalytic"):
    """BASELINE.json configs[2]: Kerr-Schild black hole (M = 1, a = 0) on the
    spherical shell of KerrSchild.yaml:80-98 (Sphere, InnerRadius 1.9, OuterRadius
    2.3, equiangular wedges, Logarithmic radial distribution, excised interior),
    DirichletAnalytic (as in the input file) or ConstraintPreservingBjorhus of type
    "ConstraintPreserving" / "ConstraintPreservingPhysical" on the outer boundary and
    DirichletAnalytic or DemandOutgoingCharSpeeds (as in EvolveGhSingleBlackHole
    set-ups, the excision surface lies inside the horizon) on the excision boundary,
    AnalyticChristoffel gauge, GaussianPlusConstant damping (:108-125)."""
    shell = domain.SphericalShell(inner_radius, outer_radius, refinement, N,
                                  radial_partitioning, radial_distribution, order=order)
    w = 11.313708499
    gam = (lambda x: analytic.gaussian_plus_constant(x, 0.001, 3.0, w),
           -1.0,
           lambda x: analytic.gaussian_plus_constant(x, 0.001, 1.0, w))
    if inner_boundary not in ("DirichletAnalytic", "DemandOutgoingCharSpeeds"):
        raise ValueError(inner_boundary)
    if outer_boundary not in ("DirichletAnalytic", "ConstraintPreserving",
                              "ConstraintPreservingPhysical"):
        raise ValueError(outer_boundary)
    ghost_dirs = {d for d, bc in ((4, inner_boundary), (5, outer_boundary))
                  if bc == "DirichletAnalytic"}
    outgoing = inner_boundary == "DemandOutgoingCharSpeeds"
    bjorhus = ((lambda g, d: outer_boundary if d == 5 else None)
               if outer_boundary != "DirichletAnalytic" else None)
    ghost = (lambda g, d: d in ghost_dirs) if ghost_dirs else False
    return Problem(lib.SYSTEM_GH, shell, lambda x, t: analytic.kerr_schild(x, mass), gam,
                   dirichlet_analytic=ghost, analytic_christoffel_gauge=True,
                   demand_outgoing=outgoing, bjorhus=bjorhus)


def gh_binary_problem(refinement, N, separation=16.0, excision_radius=0.8, object_outer_radius=4.0,
                      envelope_radius=60.0, outer_radius=300.0, opening_angle_degrees=120.0,
                      masses=(0.5, 0.5)):
    """BASELINE.json configs[4]: superposed Kerr-Schild data (synthetic) on the
    BinaryCompactObject domain (44 blocks, both objects excised; block layout of
    support/Pipelines/Bbh/Inspiral.yaml:54-101 with CubeScale 1, one N and one refinement
    level for all blocks, static maps), DirichletAnalytic with the initial data on the two
    excision spheres and the outer sphere, AnalyticChristoffel gauge of the initial data,
    constant damping parameters."""
    from . import bco
    xa, xb = 0.5 * separation, -0.5 * separation
    dom = bco.BinaryCompactObject(xa, xb, excision_radius, object_outer_radius, excision_radius,
                                  object_outer_radius, envelope_radius, outer_radius, refinement,
                                  N, opening_angle_degrees)
    centers = ((xa, 0.0, 0.0), (xb, 0.0, 0.0))
    return Problem(lib.SYSTEM_GH, dom,
                   lambda x, t: analytic.superposed_kerr_schild(x, masses, centers),
                   (1.0, -1.0, 1.0), dirichlet_analytic=True, analytic_christoffel_gauge=True)


def gh_gauge_wave_dirichlet_problem(refinement, N, amplitude=0.1, wavelength=1.0,
                                    gammas=(1.0, -1.0, 1.0)):
    """Gauge wave on the Brick [0,1]^3 that is periodic in y and z only; the x
    faces carry DirichletAnalytic with the (time-dependent) exact solution."""
    brick = domain.Brick((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), refinement, N,
                         periodic=(False, True, True))
    p = Problem(lib.SYSTEM_GH, brick,
                lambda x, t: analytic.gauge_wave(x, t, amplitude, wavelength), gammas,
                dirichlet_analytic=True)
    p.boundary_time_dependent = True
    return p


def scalar_wave_problem(refinement, N):
    """BASELINE.json configs[0] (PlaneWave3D.yaml: Brick [0,2pi]^3 periodic,
    gamma2 = 0)."""
    brick = domain.Brick([0.0] * 3, [2 * np.pi] * 3, refinement, N)
    return Problem(lib.SYSTEM_SCALAR_WAVE, brick, lambda x, t: analytic.plane_wave(x, t),
                   (0.0,))


def boundary_ghost_data(problem, part, t, halo_comps):
    """[n_external][halo_comps][N^2]: exterior state = analytic solution on the
    face, then the interior element's inverse-Jacobian row and gammas."""
    N, f = problem.N, problem.N ** 2
    C = 50 if problem.system == lib.SYSTEM_GH else 5
    hc = halo_comps
    ids = part.global_ids
    out = np.zeros((len(part.external_faces), hc, f))
    q = np.arange(f)
    a, b = q % N, q // N
    elems = sorted({le for le, _, _ in part.external_faces})
    x = dict(zip(elems, problem.coords(ids[elems])))
    J = dict(zip(elems, problem.inverse_jacobian(ids[elems])))
    S = dict(zip(elems, problem.static(ids[elems])))
    for k, (le, d, slot) in enumerate(part.external_faces):
        dim, fixed = d // 2, (N - 1 if d % 2 else 0)
        p = [fixed + N * (a + N * b), a + N * (fixed + N * b), a + N * (b + N * fixed)][dim]
        out[k, :C] = problem._initial_data(x[le][:, p], t)
        for i in range(3):
            out[k, C + i] = J[le][dim + 3 * i, p]
        if problem.system == lib.SYSTEM_GH:
            out[k, C + 3] = S[le][1, p]
            out[k, C + 4] = S[le][2, p]
        else:
            out[k, C + 3] = S[le][0, p]
    return out


class Evolution:
    """GTS evolution of one rank's share of a problem."""

    def __init__(self, problem, stepper=lib.STEPPER_ADAMS_BASHFORTH, order=3, dt=2e-4,
                 t0=0.0, gauge=lib.GAUGE_HARMONIC, gauge_params=(), device=0, world=1, rank=0,
                 process_group=None, native_exchange=True, element_order=None):
        """element_order(partition, problem) -> permutation (new -> old) of the local elements
        (single rank), e.g. sorted by step-size level for local time stepping."""
        self.world, self.rank = world, rank
        self.native_exchange = False
        self.problem = problem
        self.part = domain.Partition(problem.neighbors, world, rank,
                                     boundary_slots=problem.dirichlet_analytic,
                                     neighbor_direction=problem.orientations[0],
                                     face_permutation=problem.orientations[1],
                                     mortars=problem.mortars)
        if element_order is not None:
            self.part.reorder(element_order(self.part, problem))
        ids = self.part.global_ids
        self.ctx = lib.Context(problem.system, problem.N, self.part.n_local,
                               self.part.n_ghost, device)
        ctx = self.ctx
        if problem.bjorhus is not None:
            ln = self.part.local_neighbors
            for le, g in enumerate(ids):
                for d in range(6):
                    kind = problem.bjorhus(int(g), d) if ln[le, d] == -1 else None
                    if kind:
                        ln[le, d] = (lib.BJORHUS_PHYSICAL if kind == "ConstraintPreservingPhysical"
                                     else lib.BJORHUS)
        ctx.set_geometry(problem.inverse_jacobian(ids), problem.coords(ids),
                         self.part.local_neighbors)
        if self.part.oriented:
            ctx.set_neighbor_orientations(self.part.local_neighbor_direction,
                                          self.part.local_face_permutation)
        if len(self.part.local_mortars):
            # (sides on other ranks are ghost slots, Partition.local_mortars)
            self.local_mortars = self.part.local_mortars
            ctx.set_mortars(self.local_mortars)
        if problem.demand_outgoing:
            ctx.set_demand_outgoing_char_speeds(True)
        ctx.set_static_fields(problem.static(ids))
        if problem.system == lib.SYSTEM_GH and problem.analytic_christoffel_gauge:
            ctx.set_gauge_analytic_christoffel(problem.u0(ids, t0))
        elif problem.system == lib.SYSTEM_GH and gauge != lib.GAUGE_HARMONIC:
            ctx.set_gauge(gauge, gauge_params)
        ctx.set_state(problem.u0(ids, t0))
        if self.part.external_faces:
            ctx.set_boundary_ghost_data(self.part.n_recv, self.boundary_ghost_data(problem, t0))
        ctx.set_interior_count(self.part.n_interior)
        ctx.set_stepper(stepper, order, t0, dt)
        self.n_points = self.part.n_local * problem.N ** 3
        self._pg = process_group
        if world > 1:
            import torch
            import torch.distributed as dist
            self._torch, self._dist = torch, dist
            ctx.set_halo_map(self.part.send_map)
            f = problem.N ** 2
            per_face = ctx.halo_comps * f
            self._send = torch.as_tensor(
                _CudaArray(ctx.halo_send_ptr(), max(self.part.n_ghost, 1) * per_face),
                device=f"cuda:{device}")
            self._recv = torch.as_tensor(
                _CudaArray(ctx.halo_recv_ptr(), max(self.part.n_ghost, 1) * per_face),
                device=f"cuda:{device}")
            self._stream = torch.cuda.ExternalStream(ctx.stream, device=f"cuda:{device}")
            self._halo = HaloExchange(self.part, per_face, dist, process_group)
            if (native_exchange and dist.is_available() and dist.is_initialized()
                    and dist.get_backend(process_group) == "nccl"
                    and not problem.boundary_time_dependent):
                # the library's own communicator: rank 0 makes the ncclUniqueId
                uid = [lib.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0, group=process_group)
                ctx.comm_init(uid[0], rank, world)
                ctx.set_halo_peers(self.part.send_counts, self.part.recv_counts)
                self.native_exchange = True

    def boundary_ghost_data(self, problem, t):
        return boundary_ghost_data(problem, self.part, t, self.ctx.halo_comps)

    # -- one RHS + update ------------------------------------------------
    def _substep(self) -> bool:
        ctx = self.ctx
        t = ctx.begin_substep()
        if self.part.external_faces and self.problem.boundary_time_dependent:
            ctx.set_boundary_ghost_data(self.part.n_recv,
                                        self.boundary_ghost_data(self.problem, t))
        if self.world == 1:
            ctx.compute_time_derivative_range(t, 0, self.part.n_local)
        elif self.part.n_recv == 0:
            ctx.compute_time_derivative_range(t, 0, self.part.n_local)
        else:
            torch, dist = self._torch, self._dist
            ctx.pack_halo()
            with torch.cuda.stream(self._stream):
                works = self._halo.start(self._send, self._recv)
                if self.part.n_interior > 0:
                    ctx.compute_time_derivative_range(t, 0, self.part.n_interior)
                for w in works:
                    w.wait()
            ctx.compute_time_derivative_range(t, self.part.n_interior, self.part.n_local)
        return ctx.end_substep()

    def take_steps(self, n: int):
        if self.world == 1 or self.native_exchange:
            if not (self.part.external_faces and self.problem.boundary_time_dependent):
                self.ctx.take_steps(n)     # the whole schedule runs inside the library
                return
        done = 0
        while done < n:
            if self._substep():
                done += 1

    def gather_state(self, n_global: int):
        """Global state in global element order (rank 0 result; all ranks call)."""
        local = self.ctx.get_state()
        if self.world == 1:
            out = np.empty_like(local)
            out[self.part.global_ids] = local
            return out
        torch, dist = self._torch, self._dist
        pieces = [None] * self.world
        dist.all_gather_object(pieces, (self.part.global_ids, local), group=self._pg)
        out = np.empty((n_global,) + local.shape[1:])
        for ids, vals in pieces:
            out[ids] = vals
        return out
