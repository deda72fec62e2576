"""Front end for the reference's YAML input files (the subset of options the
accelerated path understands): tests/InputFiles/ScalarWave/PlaneWave3D.yaml,
GeneralizedHarmonic/GaugeWave3D.yaml and GeneralizedHarmonic/KerrSchild.yaml of
the reference run unchanged.

    python -m spectre_b200.input_file --input-file KerrSchild.yaml [--steps N]

The reference parses these with its Options system (src/Options/) into the
GlobalCache / DataBox of the executable named in the metadata block
(EvolveScalarWave3D, EvolveGhNoBlackHole3D, EvolveGhSingleBlackHole:
src/Evolution/Executables/); here the same option names build a
spectre_b200.evolution.Problem and drive the C-ABI.  Options outside the path
(observers' file names, AMR, phase changes, resource info) are accepted and
ignored; options that would change the numerics and are not implemented raise
InputFileError, like the reference's PARSE_ERROR.

Global time stepping with a FIXED step is what the path implements.  An input
file whose executable uses local time stepping (EvolveGhSingleBlackHole,
EvolveGhBinaryBlackHole: USE_LTS in Evolution/Executables/GeneralizedHarmonic/
CMakeLists.txt) or that lists StepChoosers (they adjust the step on every LTS
step / at every slab boundary) is therefore NOT numerically equivalent to the
reference run: it is rejected unless the caller passes
allow_gts_fixed_step=True (command line: --allow-gts-fixed-step), in which case
the ignored choosers are printed with the results.
"""
from __future__ import annotations

import argparse

import numpy as np
import yaml

from . import analytic, domain, evolution, lib


SPHERE_WEDGE_NAMES = ("UpperZ", "LowerZ", "UpperY", "LowerY", "UpperX", "LowerX")


def _expand_over_sphere_blocks(value, n_layers, what):
    """The four forms of a per-block option of the Sphere creator (Sphere.hpp:222-245,
    ExpandOverBlocks.tpp:37-120): one number, one [phi, theta, r] triple, a triple for every
    block (shells inside-out, wedges UpperZ LowerZ UpperY LowerY UpperX LowerX in each,
    Sphere.cpp:168-180), or a map from block names ("Shell0UpperZ") and block groups
    ("Shell0", "Wedges") to triples.  Returns one triple per block."""
    names = [f"Shell{l}{w}" for l in range(n_layers) for w in SPHERE_WEDGE_NAMES]

    def triple(v):
        if isinstance(v, int):
            return (v, v, v)
        if isinstance(v, (list, tuple)) and len(v) == 3 and all(isinstance(x, int) for x in v):
            return tuple(v)
        raise InputFileError(f"{what}: expected a number or a list of three numbers, got {v!r}")
    if isinstance(value, dict):
        groups = {"Wedges": names}
        for l in range(n_layers):
            groups[f"Shell{l}"] = names[6 * l:6 * l + 6]
        per_block = {}
        for key, v in value.items():
            members = groups.get(key, [key])
            for name in members:
                if name not in names:
                    raise InputFileError(f"{what}: unknown block or group '{key}'")
                if name in per_block:
                    raise InputFileError(f"{what}: duplicate block name '{name}' "
                                         f"(expanded from '{key}')")
                per_block[name] = triple(v)
        missing = [n for n in names if n not in per_block]
        if missing:
            raise InputFileError(f"{what}: value for block '{missing[0]}' is missing")
        return [per_block[n] for n in names]
    if isinstance(value, (list, tuple)) and value and isinstance(value[0], (list, tuple)):
        if len(value) != len(names):
            raise InputFileError(f"{what}: you supplied {len(value)} values, but the domain "
                                 f"creator has {len(names)} blocks")
        return [triple(v) for v in value]
    return [triple(value)] * len(names)


class InputFileError(ValueError):
    pass


def _one(options, what):
    """A YAML map with exactly one key (the reference's factory-creatable idiom)."""
    if isinstance(options, str):
        return options, {}
    if not isinstance(options, dict) or len(options) != 1:
        raise InputFileError(f"{what}: expected exactly one option group, got {options!r}")
    (name, body), = options.items()
    return name, (body or {})


def _gaussian_plus_constant(opts, what):
    name, o = _one(opts, what)
    if name == "Constant":
        value = float(o["Value"])
        return value
    if name != "GaussianPlusConstant":
        raise InputFileError(f"{what}: damping function {name} is not implemented")
    c, a, w = float(o["Constant"]), float(o["Amplitude"]), float(o["Width"])
    center = tuple(float(v) for v in o["Center"])
    if a == 0.0:
        return c
    return lambda x: analytic.gaussian_plus_constant(x, c, a, w, center)


STEPPERS = {
    "AdamsBashforth": lib.STEPPER_ADAMS_BASHFORTH, "Rk3HesthavenSsp": lib.STEPPER_RK3_HESTHAVEN,
    "Rk3Owren": lib.STEPPER_RK3_OWREN, "Rk3Kennedy": lib.STEPPER_RK3_KENNEDY,
    "ClassicalRungeKutta4": lib.STEPPER_RK4, "DormandPrince5": lib.STEPPER_DORMAND_PRINCE5,
}


class Run:
    """Everything an input file determines for the path."""

    LTS_EXECUTABLES = ("EvolveGhSingleBlackHole", "EvolveGhBinaryBlackHole")

    def __init__(self, metadata, options, allow_gts_fixed_step=False, lts_fixed_levels=False):
        self.metadata, self.options = metadata or {}, options
        exe = str(self.metadata.get("Executable", ""))
        self.system = lib.SYSTEM_SCALAR_WAVE if "ScalarWave" in exe else lib.SYSTEM_GH
        ev = options["Evolution"]
        self.t0, self.dt = float(ev["InitialTime"]), float(ev["InitialTimeStep"])
        # a slab is InitialSlabSize long (default: one step); triggers count slabs
        self.steps_per_slab = int(round(float(ev.get("InitialSlabSize", self.dt)) / self.dt))
        name, o = _one(ev["TimeStepper"], "TimeStepper")
        if name not in STEPPERS:
            raise InputFileError(f"TimeStepper {name} is not implemented")
        self.stepper, self.order = STEPPERS[name], int(o.get("Order", 0))
        self.step_choosers_ignored = []
        if "StepChoosers" in ev and ev["StepChoosers"]:
            self.step_choosers_ignored = [list(c)[0] if isinstance(c, dict) else str(c)
                                          for c in ev["StepChoosers"]]
        self.lts_executable = any(exe.startswith(name) for name in self.LTS_EXECUTABLES)
        # local time stepping with the step sizes fixed at the start (run_lts): what the
        # reference's LTS executable starts with -- the largest step slab / 2^n below
        # InitialTimeStep and, if listed, below the ElementSizeCfl goal -- without the
        # later step-size changes of the choosers (LimitIncrease, ErrorControl, ...)
        self.lts_fixed_levels = bool(lts_fixed_levels)
        self.slab_size = float(ev.get("InitialSlabSize", self.dt))
        self.element_size_cfl = None
        for ch in ev.get("StepChoosers") or []:
            if isinstance(ch, dict) and "ElementSizeCfl" in ch:
                self.element_size_cfl = float(ch["ElementSizeCfl"]["SafetyFactor"])
        if self.lts_fixed_levels:
            if name != "AdamsBashforth":
                raise InputFileError("local time stepping is implemented for AdamsBashforth")
            self.step_choosers_ignored = [c for c in self.step_choosers_ignored
                                          if c != "ElementSizeCfl"]
        if (self.lts_executable or self.step_choosers_ignored) and not allow_gts_fixed_step \
                and not self.lts_fixed_levels:
            why = []
            if self.lts_executable:
                why.append(f"executable {exe} uses local time stepping")
            if self.step_choosers_ignored:
                why.append("StepChoosers " + ", ".join(self.step_choosers_ignored)
                           + " would change the step size")
            raise InputFileError(
                "; ".join(why) + ": the path takes fixed global time steps, so the run would not "
                "be numerically equivalent to the reference executable (pass "
                "allow_gts_fixed_step=True / --allow-gts-fixed-step to run it anyway)")
        sd = options["SpatialDiscretization"]
        bc_name, _ = _one(sd["BoundaryCorrection"], "BoundaryCorrection")
        if bc_name != "UpwindPenalty":
            raise InputFileError(f"BoundaryCorrection {bc_name} is not implemented")
        dgo = sd["DiscontinuousGalerkin"]
        if dgo.get("Formulation", "StrongInertial") != "StrongInertial" or \
                dgo.get("Quadrature", "GaussLobatto") != "GaussLobatto":
            raise InputFileError("only StrongInertial on GaussLobatto points is implemented")
        self.filter = None
        flt = (options.get("Filtering") or {}).get("ExpFilter0")
        if flt and flt.get("Enable", True):
            self.filter = (float(flt["Alpha"]), int(flt["HalfPower"]))
        self._initial_data()
        self._domain()
        self._system()
        self._events()

    # -- InitialData ---------------------------------------------------------
    def _initial_data(self):
        name, o = _one(self.options["InitialData"], "InitialData")
        self.initial_data_name = name
        if name == "PlaneWave":
            pname, prof = _one(o["Profile"], "PlaneWave.Profile")
            if pname != "Sinusoid":
                raise InputFileError(f"PlaneWave profile {pname} is not implemented")
            k, c = tuple(map(float, o["WaveVector"])), tuple(map(float, o["Center"]))
            A, wn, ph = float(prof["Amplitude"]), float(prof["Wavenumber"]), float(prof["Phase"])
            self.u0 = lambda x, t: analytic.plane_wave(x, t, k, c, A, wn, ph)
            self.time_dependent = True
        elif name == "GeneralizedHarmonic(GaugeWave)":
            A, wl = float(o["Amplitude"]), float(o["Wavelength"])
            self.gauge_wave = (A, wl)
            self.u0 = lambda x, t: analytic.gauge_wave(x, t, A, wl)
            self.time_dependent = True
        elif name == "GeneralizedHarmonic(KerrSchild)":
            if any(float(v) != 0.0 for v in list(o["Spin"]) + list(o["Velocity"])):
                raise InputFileError("KerrSchild with spin or boost is not implemented")
            M, c = float(o["Mass"]), tuple(map(float, o["Center"]))
            self.u0 = lambda x, t: analytic.kerr_schild(x, M, c)
            self.time_dependent = False
        else:
            raise InputFileError(f"InitialData {name} is not implemented")

    # -- DomainCreator -------------------------------------------------------
    def _boundary_condition(self, opts, what):
        name, o = _one(opts, what)
        if name == "DirichletAnalytic":
            return "DirichletAnalytic"
        if name == "DemandOutgoingCharSpeeds":
            return "DemandOutgoingCharSpeeds"
        if name == "ConstraintPreservingBjorhus":
            return str(o["Type"])
        if name == "Periodic":
            return "Periodic"
        raise InputFileError(f"{what}: boundary condition {name} is not implemented")

    def _domain(self):
        name, o = _one(self.options["DomainCreator"], "DomainCreator")
        if o.get("TimeDependence", None) not in (None, "None") or \
                o.get("TimeDependentMaps", None) not in (None, "None"):
            raise InputFileError("time-dependent maps are not implemented")
        self.ghost, self.outgoing, self.bjorhus = False, False, None
        if name == "Brick":
            pts = o["InitialGridPoints"]
            if len(set(pts)) != 1:
                raise InputFileError("anisotropic InitialGridPoints are not implemented")
            bcs = o["BoundaryConditions"]
            kinds = []
            for b in bcs:
                if isinstance(b, dict) and set(b) == {"Lower", "Upper"}:
                    kinds.append((self._boundary_condition(b["Lower"], "Brick.Lower"),
                                  self._boundary_condition(b["Upper"], "Brick.Upper")))
                else:
                    k = self._boundary_condition(b, "Brick.BoundaryConditions")
                    kinds.append((k, k))
            for dim, k in enumerate(kinds):
                if (k[0] == "Periodic") != (k[1] == "Periodic"):
                    raise InputFileError(
                        f"Brick: periodic boundary condition on only one side of dimension {dim} "
                        "(both or neither must be Periodic)")
            periodic = tuple(k == ("Periodic", "Periodic") for k in kinds)
            self.domain = domain.Brick(o["LowerBound"], o["UpperBound"], o["InitialRefinement"],
                                       int(pts[0]), periodic=periodic)
            face_bc = {2 * dim + side: kinds[dim][side] for dim in range(3) for side in range(2)
                       if not periodic[dim]}
        elif name == "Sphere":
            iname, io = _one(o["Interior"], "Sphere.Interior")
            if iname != "ExciseWithBoundaryCondition":
                raise InputFileError("Sphere with a filled interior is not implemented")
            if o.get("WhichWedges", "All") != "All" or \
                    o.get("EquatorialCompression", None) not in (None, "None"):
                raise InputFileError("Sphere: WhichWedges / EquatorialCompression not implemented")
            partitioning = tuple(map(float, o.get("RadialPartitioning", [])))
            n_layers = len(partitioning) + 1
            pts_blocks = _expand_over_sphere_blocks(o["InitialGridPoints"], n_layers,
                                                    "Sphere.InitialGridPoints")
            if len({p for blk in pts_blocks for p in blk}) != 1:
                raise InputFileError("InitialGridPoints that differ between blocks or dimensions "
                                     "(p-refinement, anisotropic meshes) are not implemented")
            pts = pts_blocks[0][0]
            ref_blocks = _expand_over_sphere_blocks(o["InitialRefinement"], n_layers,
                                                    "Sphere.InitialRefinement")
            if any(r[0] != r[1] for r in ref_blocks):
                raise InputFileError("different angular refinement levels in one block are not "
                                     "implemented")
            ref = [[(r[0], r[2]) for r in ref_blocks[6 * l:6 * l + 6]] for l in range(n_layers)]
            ref = [layer[0] if len(set(layer)) == 1 else layer for layer in ref]
            self.domain = domain.SphericalShell(
                float(o["InnerRadius"]), float(o["OuterRadius"]), ref, int(pts),
                partitioning,
                list(o.get("RadialDistribution", ["Linear"])),
                bool(o.get("UseEquiangularMap", True)))
            face_bc = {4: self._boundary_condition(io, "Sphere.Interior"),
                       5: self._boundary_condition(o["OuterBoundaryCondition"],
                                                   "Sphere.OuterBoundaryCondition")}
        elif name == "BinaryCompactObject":
            self._binary_compact_object(o)
            return
        else:
            raise InputFileError(f"DomainCreator {name} is not implemented")
        self.face_bc = face_bc
        ghost_dirs = {d for d, k in face_bc.items() if k == "DirichletAnalytic"}
        if ghost_dirs:
            self.ghost = lambda g, d: d in ghost_dirs
        self.outgoing = any(k == "DemandOutgoingCharSpeeds" for k in face_bc.values())
        bj = {d: k for d, k in face_bc.items() if k.startswith("ConstraintPreserving")}
        if bj:
            if len({d // 2 for d in bj}) > 1:
                # the reference applies the external faces of an element one after the other,
                # each seeing the dt corrected by the previous ones on shared edge points
                # (BoundaryConditionsImpl.hpp:277-278, 636-660); the path evaluates every
                # Bjorhus face from the uncorrected volume dt
                raise InputFileError(
                    "ConstraintPreservingBjorhus in more than one dimension (elements with two "
                    "or more Bjorhus faces) is not implemented")
            self.bjorhus = lambda g, d: bj.get(d)

    def _binary_compact_object(self, o):
        """domain::creators::BinaryCompactObject (BinaryCompactObject.hpp option list): the
        subset spectre_b200.bco builds -- both objects excised, CubeScale 1, no centre-of-mass
        offset, static maps, one number of grid points for all blocks and dimensions."""
        from . import bco
        if float(o.get("CubeScale", 1.0)) != 1.0:
            raise InputFileError("BinaryCompactObject: CubeScale != 1 (focally offset wedges) is "
                                 "not implemented")
        if any(float(v) != 0.0 for v in o.get("CenterOfMassOffset", [0.0, 0.0])):
            raise InputFileError("BinaryCompactObject: CenterOfMassOffset is not implemented")
        objects = {}
        for tag in ("ObjectA", "ObjectB"):
            ob = o[tag]
            if "InnerRadius" not in ob:
                raise InputFileError(f"BinaryCompactObject: {tag} without an excised sphere "
                                     "(CartesianCubeAtXCoord) is not implemented")
            iname, io = _one(ob["Interior"], f"{tag}.Interior")
            if iname != "ExciseWithBoundaryCondition":
                raise InputFileError(f"BinaryCompactObject: {tag} with a filled interior is not "
                                     "implemented")
            objects[tag] = (float(ob["XCoord"]), float(ob["InnerRadius"]),
                            float(ob["OuterRadius"]), bool(ob.get("UseLogarithmicMap", False)),
                            self._boundary_condition(io, f"{tag}.Interior"))
        if objects["ObjectA"][3] != objects["ObjectB"][3]:
            raise InputFileError("BinaryCompactObject: different UseLogarithmicMap for the two "
                                 "objects is not implemented")
        groups = bco.BinaryCompactObject.GROUPS

        def per_group(value, what, length):
            if isinstance(value, dict):
                missing = [g for g in groups if g not in value]
                if missing:
                    raise InputFileError(f"{what}: no entry for {missing}")
                out = {g: value[g] for g in groups}
            else:
                out = {g: value for g in groups}
            return {g: (tuple(int(x) for x in v) if isinstance(v, (list, tuple))
                        else (int(v),) * length) for g, v in out.items()}
        pts = per_group(o["InitialGridPoints"], "BinaryCompactObject.InitialGridPoints", 3)
        if len({p for v in pts.values() for p in v}) != 1:
            raise InputFileError("InitialGridPoints that differ between blocks or dimensions "
                                 "(p-refinement, anisotropic meshes) are not implemented")
        ref = per_group(o["InitialRefinement"], "BinaryCompactObject.InitialRefinement", 3)
        env, shell = o["Envelope"], o["OuterShell"]
        (xa, ra_in, ra_out, log_a, bc_a), (xb, rb_in, rb_out, _, bc_b) = (objects["ObjectA"],
                                                                          objects["ObjectB"])
        try:
            self.domain = bco.BinaryCompactObject(
                xa, xb, ra_in, ra_out, rb_in, rb_out, float(env["Radius"]), float(shell["Radius"]),
                ref, pts["Envelope"][0], float(shell.get("OpeningAngle", 90.0)),
                bool(o.get("UseEquiangularMap", True)), log_a,
                str(env.get("RadialDistribution", "Linear")),
                str(shell.get("RadialDistribution", "Linear")))
        except ValueError as err:
            raise InputFileError(f"BinaryCompactObject: {err}") from None
        kinds = {"excision_a": bc_a, "excision_b": bc_b,
                 "outer": self._boundary_condition(shell["BoundaryCondition"],
                                                   "OuterShell.BoundaryCondition")}
        dom = self.domain
        kind_of = lambda g, d: kinds[dom.external_boundary(g, d)]   # noqa: E731
        self.face_bc = kinds
        if "DirichletAnalytic" in kinds.values():
            self.ghost = lambda g, d: kind_of(g, d) == "DirichletAnalytic"
        self.outgoing = "DemandOutgoingCharSpeeds" in kinds.values()
        if any(k.startswith("ConstraintPreserving") for k in kinds.values()):
            self.bjorhus = lambda g, d: (kind_of(g, d)
                                         if kind_of(g, d).startswith("ConstraintPreserving")
                                         else None)

    # -- EvolutionSystem -----------------------------------------------------
    def _system(self):
        self.gauge, self.gauge_params, self.analytic_christoffel = lib.GAUGE_HARMONIC, (), False
        if self.system == lib.SYSTEM_SCALAR_WAVE:
            self.static = (0.0,)     # gamma2 = 0 in the executable (ScalarWave/Initialize.hpp:48-49)
            return
        gh = self.options["EvolutionSystem"]["GeneralizedHarmonic"]
        gname, go = _one(gh["GaugeCondition"], "GaugeCondition")
        if gname == "Harmonic":
            pass
        elif gname == "AnalyticChristoffel":
            if self.initial_data_name == "GeneralizedHarmonic(GaugeWave)":
                self.gauge = lib.GAUGE_ANALYTIC_GAUGE_WAVE
                self.gauge_params = self.gauge_wave
            else:
                self.analytic_christoffel = True
        elif gname == "DampedHarmonic":
            self.gauge = lib.GAUGE_DAMPED_HARMONIC
            self.gauge_params = (float(go["SpatialDecayWidth"]), *map(float, go["Amplitudes"]),
                                 *map(int, go["Exponents"]))
        else:
            raise InputFileError(f"GaugeCondition {gname} is not implemented")
        self.static = tuple(_gaussian_plus_constant(gh[f"DampingFunctionGamma{i}"],
                                                    f"DampingFunctionGamma{i}") for i in range(3))

    # -- EventsAndTriggers ---------------------------------------------------
    def _events(self):
        self.n_steps, self.observe_interval = None, None
        for et in self.options.get("EventsAndTriggers") or []:
            trig = et.get("Trigger")
            events = [list(e)[0] if isinstance(e, dict) else e for e in et.get("Events", [])]
            if "Completion" in events and isinstance(trig, dict):
                tname, to = _one(trig, "Trigger")
                spec = (to or {}).get("Specified", {}).get("Values")
                if tname == "Slabs" and spec:
                    self.n_steps = int(spec[0]) * self.steps_per_slab
                elif tname == "Times" and spec:
                    self.n_steps = int(round((float(spec[0]) - self.t0) / self.dt))
            if "ObserveNorms" in events and isinstance(trig, dict):
                tname, to = _one(trig, "Trigger")
                if tname == "Slabs" and "EvenlySpaced" in (to or {}):
                    self.observe_interval = int(to["EvenlySpaced"]["Interval"]) * \
                        self.steps_per_slab

    # -- build + run ---------------------------------------------------------
    def problem(self):
        p = evolution.Problem(self.system, self.domain, self.u0, self.static,
                              dirichlet_analytic=self.ghost,
                              analytic_christoffel_gauge=self.analytic_christoffel,
                              demand_outgoing=self.outgoing, bjorhus=self.bjorhus)
        p.boundary_time_dependent = bool(self.ghost) and self.time_dependent
        return p

    def evolution(self, device=0, world=1, rank=0, process_group=None):
        ev = evolution.Evolution(self.problem(), self.stepper, self.order, self.dt, self.t0,
                                 self.gauge, self.gauge_params, device, world, rank,
                                 process_group)
        if self.filter:
            ev.ctx.set_exponential_filter(True, *self.filter)
        return ev

    def run(self, n_steps=None, device=0, observe=None):
        """Evolve and return [(step, time, {name: L2 error norm})] -- ObserveNorms of
        Error(...) with NormType L2Norm, Components Sum
        (ParallelAlgorithms/Events/ObserveNorms.hpp:60-80)."""
        n_steps = self.n_steps if n_steps is None else n_steps
        if n_steps is None:
            raise InputFileError("no Completion trigger with a specified slab or time")
        observe = observe or self.observe_interval or n_steps
        ev = self.evolution(device)
        names = ("Psi", "Pi", "Phi") if self.system == lib.SYSTEM_SCALAR_WAVE else \
            ("SpacetimeMetric", "Pi", "Phi")
        blocks = ((0, 1), (1, 2), (2, 5)) if self.system == lib.SYSTEM_SCALAR_WAVE else \
            ((0, 10), (10, 20), (20, 50))
        out, done = [], 0
        ids = ev.part.global_ids

        def observe_now():
            state = ev.ctx.get_state()
            exact = ev.problem.u0(ids, ev.ctx.time)
            npts = state.shape[0] * state.shape[2]
            norms = {f"Error({nm})": float(np.sqrt(np.sum((state[:, a:b] - exact[:, a:b]) ** 2)
                                                   / npts)) for nm, (a, b) in zip(names, blocks)}
            out.append((done, ev.ctx.time, norms))
        observe_now()
        while done < n_steps:
            k = min(observe - done % observe, n_steps - done)
            ev.take_steps(k)
            done += k
            observe_now()
        if self.outgoing:
            ev.ctx.check_outgoing_char_speeds()
        ev.ctx.close()
        return out


    def run_lts(self, n_slabs=None, device=0):
        """Evolve with Adams-Bashforth local time stepping, the element steps fixed at the
        start (see __init__), and return [(slab, time, {name: L2 error norm})] observed at the
        slab boundaries (all elements are at the same time there)."""
        from . import lts
        if n_slabs is None:
            if self.n_steps is None:
                raise InputFileError("no Completion trigger with a specified slab or time")
            n_slabs = max(1, self.n_steps // self.steps_per_slab)
        if self.gauge not in (lib.GAUGE_HARMONIC, lib.GAUGE_DAMPED_HARMONIC):
            raise InputFileError("local time stepping: this gauge is not implemented")
        problem = self.problem()
        n_el = problem.brick.n_elements
        ev = lts.LtsEvolution(
            problem, self.order, self.slab_size, self.t0,
            step_goal=None if self.element_size_cfl is not None else np.full(n_el, self.dt),
            safety_factor=self.element_size_cfl or 1.0, max_step=self.dt, past="analytic",
            gauge=self.gauge, gauge_params=self.gauge_params, device=device,
            filter_params=self.filter)
        names = ("Psi", "Pi", "Phi") if self.system == lib.SYSTEM_SCALAR_WAVE else \
            ("SpacetimeMetric", "Pi", "Phi")
        blocks = ((0, 1), (1, 2), (2, 5)) if self.system == lib.SYSTEM_SCALAR_WAVE else \
            ((0, 10), (10, 20), (20, 50))
        out = []
        per_slab = int(round(self.slab_size / ev.dt_coarse))

        def observe_now(k):
            ids, state = ev.state()
            exact = problem.u0(ids, ev.time)
            npts = state.shape[0] * state.shape[2]
            norms = {f"Error({nm})": float(np.sqrt(np.sum((state[:, a:b] - exact[:, a:b]) ** 2)
                                                   / npts)) for nm, (a, b) in zip(names, blocks)}
            out.append((k, ev.time, norms))
        observe_now(0)
        for k in range(n_slabs):
            ev.take_coarse_steps(per_slab)
            observe_now(k + 1)
        if self.outgoing:
            ev.ctx.check_outgoing_char_speeds()
        self.lts_levels = ev.levels
        self.lts_dt_coarse = ev.dt_coarse
        ev.ctx.close()
        return out


def load(path, allow_gts_fixed_step=False, lts_fixed_levels=False):
    with open(path) as f:
        docs = list(yaml.safe_load_all(f))
    if len(docs) == 1:
        return Run({}, docs[0], allow_gts_fixed_step, lts_fixed_levels)
    return Run(docs[0], docs[1], allow_gts_fixed_step, lts_fixed_levels)


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--input-file", required=True)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--allow-gts-fixed-step", action="store_true",
                    help="run an LTS / step-chooser input file with fixed global steps (NOT "
                         "numerically equivalent to the reference executable)")
    ap.add_argument("--lts-fixed-levels", action="store_true",
                    help="run with Adams-Bashforth local time stepping, the element steps fixed "
                         "at what InitialTimeStep / ElementSizeCfl give at the start (the other "
                         "step choosers are ignored and printed)")
    args = ap.parse_args()
    run = load(args.input_file, args.allow_gts_fixed_step, args.lts_fixed_levels)
    if args.lts_fixed_levels:
        print(f"local time stepping with fixed steps; ignored StepChoosers: "
              f"{run.step_choosers_ignored or 'none'} -- the reference's run changes its steps "
              "with them")
        for slab, t, norms in run.run_lts():
            print(f"slab {slab:6d}  t = {t:.6f}  " +
                  "  ".join(f"{k} = {v:.6e}" for k, v in norms.items()))
        print("steps:", {int(l): f"{run.lts_dt_coarse / 2 ** int(l):.6g}"
                         for l in sorted(set(run.lts_levels.tolist()))})
        return
    if run.lts_executable or run.step_choosers_ignored:
        print("WARNING: fixed global time steps; ignored StepChoosers: "
              f"{run.step_choosers_ignored or 'none'}; LTS executable: {run.lts_executable} "
              "-- not numerically equivalent to the reference run")
    for step, t, norms in run.run(args.steps):
        print(f"step {step:6d}  t = {t:.6f}  " + "  ".join(f"{k} = {v:.6e}" for k, v in norms.items()))


if __name__ == "__main__":
    main()
