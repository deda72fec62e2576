#!/usr/bin/env python
"""bench.py -- fp64 DG grid-point RHS updates/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # CPU arm (the oracle port)

Workload (config.workload), default = the north-star configuration, BASELINE.json
configs[3]: GeneralizedHarmonic Kerr-Schild black hole (M = 1, a = 0) on the
Sphere domain with excision, six equiangular wedges x 4^3 angular x 2^4 radial
= 6144 elements per GPU, N = P+1 = 12 Legendre-Gauss-Lobatto points per
dimension (10.6 M grid points, 4.25 GB per state copy per GPU),
DirichletAnalytic boundaries, AnalyticChristoffel gauge, GaussianPlusConstant
damping, exponential filter, Adams-Bashforth 3 -- tests/InputFiles/
GeneralizedHarmonic/KerrSchild.yaml:73-132 at the resolution of configs[3].
A "step" is one full AB3 time step of all elements = one RHS evaluation per
grid point + the stepper update + the filter.  Multi-GPU runs are weak scaling:
the shell grows outwards by radial layers and is cut at constant radius, the
cut mortar faces are exchanged over NCCL.  `--workload gauge-wave` is
BASELINE.json configs[1] (GH gauge wave, 16^3 elements, N = 8), also measured
in every default run and reported as `secondary` in the same JSON line.

The JSON line carries: value (state resident in HBM), e2e (state copied
host->device and back every step through the C-ABI), roofline of the dominant
kernel (CUDA-event timed inside this run), cpu_baseline (oracle port on the
host cores, bounded sample of the same workload), clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 DG grid-point RHS updates/sec"
UNIT = "grid-point-updates/s"

WORKLOAD_DEFAULTS = {            # refine, points per dimension, CPU-sample refine
    "kerr-schild-shell": (3, 12, 1),
    "gauge-wave": (4, 8, 3),
    "kerr-schild": (4, 12, 2),
    "bbh": (2, 12, 0),
}
FILTER = (36.0, 64)              # KerrSchild.yaml:127-132 (Alpha, HalfPower)


def b_alg(N, n_u=50, G=3, T_f=6, k=3):
    """Algorithmic bytes per grid-point update, SURVEY.md 8(d)."""
    return 8.0 * ((2 * n_u + 9 + G) + (k + 2) * n_u + (6.0 / N) * (2 * (n_u + T_f) + 2 * n_u))


def kernel_alg_bytes(N, n_u=50, n_static=3, k=3):
    """Compulsory bytes per grid point of each of OUR kernels:
    face: both sides' face values + 5 face statics in, lifted corrections out;
    volume: u, J^-1, statics, corrections in, dt_u out; update: u, k derivs in,
    u out; filter: u in, u out."""
    face = 8.0 * (6.0 / N) * (2 * (n_u + 5) + n_u)
    volume = 8.0 * (2 * n_u + 9 + n_static + (6.0 / N) * n_u)
    update = 8.0 * (k + 2) * n_u
    return {"face": face, "volume": volume, "update": update, "filter": 8.0 * 2 * n_u}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons DURING the timed region, read through
    NVML in this process (nvidia_ml_py), a few samples only (every 250 ms).  Every query
    perturbs a multi-GPU run: measured on the 2-GPU box, `nvidia-smi -lms 20` stalls NCCL
    progress for ~65 ms per query (configs[1] at 2 GPUs: 1.65 -> 4.65 ms per step), and
    NVML polled every 50 ms still costs 4-20 % (1.68 -> 1.75 ms; shell N = 12: 13.5 ->
    16.2 ms per step)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, device_index, period_s=0.25):
        self.idx, self.period = device_index, period_s
        self.samples, self.bits = [], 0
        self.smax = None
        self._stop = threading.Event()
        self.thread = None
        self.error = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # no NVML: report that instead of a number
            self.error = f"NVML unavailable: {e}"
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(reasons(self.h))
            except Exception as e:
                self.error = str(e)
                return
            self._stop.wait(self.period)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.error or "no sampler"]}
        self._stop.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.smax,
                "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b),
                "samples": len(self.samples), "how": "NVML in-process, 250 ms period"}


def weak_refinement(world, base):
    """base^3 elements per rank: double one dimension per factor of two."""
    ref = [base, base, base]
    w, d = world, 0
    while w > 1:
        ref[d % 3] += 1
        w //= 2
        d += 1
    return ref


def shell_radial_level(args, world):
    """Radial refinement level of the shell workload: 2^(refine+1) radial elements
    on one GPU, doubled with the GPU count (weak scaling); strong scaling: fixed
    2^(refine+1) * strong_factor radial elements."""
    if getattr(args, "scaling", "weak") == "strong":
        return args.refine + 1 + int(np.log2(args.strong_factor))
    lr, w = args.refine + 1, world
    while w > 1:
        lr += 1
        w //= 2
    return lr


def shell_problem(evolution, refine, lr, N, inner_boundary="DirichletAnalytic",
                  outer_boundary="DirichletAnalytic"):
    # radial element ratio q such that elements are a quarter as deep as they are
    # wide at every radius; the outer radius grows with the number of radial elements
    q = 1.0 + 0.125 * np.pi / 2 ** refine
    return evolution.gh_kerr_schild_shell_problem(
        (refine, lr), N, inner_radius=1.9, outer_radius=1.9 * q ** (2 ** lr),
        order="radial", inner_boundary=inner_boundary, outer_boundary=outer_boundary)


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample of the workload).
# The Charm++ executables of the reference cannot be built here,
# so this arm times the CPU restatement of the same path: oracle/dg_oracle.c
# built -O3 -march=native, OpenMP over elements for the RHS, the AB3 update
# (orc_lincomb) and the filter (orc_apply_filter), with every host thread.
# ---------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuArm:
    def __init__(self, workload, N, sample_refine, dt, use_filter):
        import ctypes
        from oracle import oracle as orc
        self.orc, self.ct = orc, ctypes
        self.L = orc.use_optimized_build()   # timing build (never used for parity)
        # torchrun exports OMP_NUM_THREADS=1 to its workers: set the count ourselves
        self.L.orc_set_num_threads(host_threads())
        self.threads = int(self.L.orc_num_threads())
        self.N, self.dt = N, dt
        if workload in ("kerr-schild-shell", "bbh"):
            from spectre_b200 import domain, evolution
            problem = (shell_problem(evolution, sample_refine, sample_refine + 1, N)
                       if workload == "kerr-schild-shell"
                       else evolution.gh_binary_problem(sample_refine, N))
            part = domain.Partition(problem.neighbors, 1, 0,
                                    boundary_slots=problem.dirichlet_analytic,
                                    neighbor_direction=problem.orientations[0],
                                    face_permutation=problem.orientations[1],
                                    mortars=problem.mortars)
            ids = part.global_ids
            x, J = problem.coords(ids), problem.inverse_jacobian(ids)
            u0 = problem.u0(ids, 0.0)
            H = np.zeros((len(ids), 4, N ** 3))
            dH = np.zeros((len(ids), 16, N ** 3))
            for e in range(len(ids)):
                H[e], dH[e] = orc.analytic_christoffel_gauge(N, u0[e], J[e])
            sf = np.concatenate([problem.static(ids), H, dH], axis=1)
            ext = evolution.boundary_ghost_data(problem, part, 0.0, 55)[:, :50]
            nbr, nd, perm = (part.local_neighbors, part.local_neighbor_direction,
                             part.local_face_permutation)
            self.rhs = lambda v: orc.dg_rhs(1, N, v, J, sf, nbr, gauge_params=orc.GAUGE_GIVEN,
                                            ext_u=ext, nbr_dir=nd, face_perm=perm)
            self.u = np.ascontiguousarray(u0)
            self.desc = (f"Kerr-Schild shell, refinement ({sample_refine}, {sample_refine + 1}): "
                         f"{len(ids)} elements" if workload == "kerr-schild-shell" else
                         f"BinaryCompactObject domain, refinement {sample_refine}: "
                         f"{len(ids)} elements")
        else:
            b = orc.Brick([0, 0, 0], [1, 1, 1], [sample_refine] * 3, N)
            x, J, nb = b.coords(), b.inverse_jacobian(), b.neighbors()
            u0 = np.stack([orc.gh_vars_from_metric(*orc.gauge_wave_metric(x[e], 0.0))
                           for e in range(b.nelem)])
            stat = np.zeros((b.nelem, 3, b.n))
            stat[:, 0], stat[:, 1], stat[:, 2] = 1.0, -1.0, 1.0
            self.rhs = lambda v: orc.dg_rhs(1, N, v, J, stat, nb)
            self.u = np.ascontiguousarray(u0)
            self.desc = f"gauge wave, refinement {sample_refine}: {b.nelem} elements"
        self.F = (np.ascontiguousarray(orc.exponential_filter_matrix(
            N, *((36.0, 420) if workload == "bbh" else FILTER))) if use_filter else None)
        self.points = self.u.shape[0] * N ** 3
        self.hist = [self.rhs(self.u) for _ in range(2)]   # AB3 history (static start)

    def step(self):
        ct, orc = self.ct, self.orc
        self.hist.append(self.rhs(self.u))
        dt = self.dt
        coefs = np.array([5.0 / 12.0 * dt, -4.0 / 3.0 * dt, 23.0 / 12.0 * dt])
        ptrs = (ct.c_void_p * 3)(*[h.ctypes.data for h in self.hist])
        self.L.orc_lincomb(ct.c_longlong(self.u.size), ct.c_double(1.0), orc._p(self.u), 3,
                           orc._p(coefs), ptrs)
        if self.F is not None:
            self.L.orc_apply_filter(self.N, ct.c_longlong(self.u.shape[0] * self.u.shape[1]),
                                    orc._p(self.F), orc._p(self.u))
        self.hist.pop(0)

    def run(self, steps, warmup, budget_s):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            self.step()
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        el = time.perf_counter() - t0
        assert np.isfinite(self.u).all()
        v = self.points * done / el
        return {"value": v, "unit": UNIT, "cores": self.threads, "kind": "port",
                "value_per_core": v / self.threads,
                "sample": f"{self.desc}, N={self.N}, {done} AB3 steps"
                          f"{' + filter' if self.F is not None else ''} in {el:.1f} s; "
                          "oracle/dg_oracle.c (gcc -O3 -march=native), OpenMP over elements "
                          f"for RHS, update and filter, {self.threads} threads",
                "steps": done, "ms_per_step": 1e3 * el / done}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(args.workload, args.points, args.cpu_sample_refine, args.dt,
                 args.workload in ("kerr-schild-shell", "bbh") and not args.no_filter)
    r = arm.run(args.steps, min(args.warmup, 1), budget_s=60.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 1),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1, cpu=True),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample",
                                           "value_per_core")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "reference Charm++ executable cannot be built here (no Charm++/Blaze/...): "
                "this arm times the CPU restatement (oracle) of the same path on a bounded "
                "sample of the workload (config.cpu_sample)",
    }
    print(json.dumps(line))


WORKLOAD_NAMES = {
    "gauge-wave": "BASELINE.json configs[1]: GeneralizedHarmonic gauge wave (A=0.1, "
                  "lambda=1), periodic Brick [0,1]^3 per GPU (the domain grows by whole "
                  "wavelengths with the GPU count), AB3, dt=2e-4, UpwindPenalty, "
                  "gamma0/1/2=1/-1/1",
    "kerr-schild": "BASELINE.json configs[2]/[3] stand-in: GeneralizedHarmonic Kerr-Schild "
                   "(M=1, a=0) on a Brick lattice (elements of edge M/8 from x=2M), "
                   "DirichletAnalytic boundaries, AnalyticChristoffel gauge, "
                   "GaussianPlusConstant damping (KerrSchild.yaml), AB3, dt=2e-4",
    "kerr-schild-shell": "BASELINE.json configs[3] (configs[2] geometry at P=11): "
                         "GeneralizedHarmonic Kerr-Schild (M=1, a=0) on the Sphere domain with "
                         "excision (six equiangular wedges per layer, Logarithmic radial "
                         "distribution, inner radius 1.9 M, h-refined: 4^L angular x 2^Lr radial "
                         "elements per wedge; weak scaling adds radial layers outwards and cuts "
                         "the shell at constant radius), DirichletAnalytic boundaries, "
                         "AnalyticChristoffel gauge, GaussianPlusConstant damping, exponential "
                         "filter (KerrSchild.yaml:73-132), AB3, dt=2e-4",
    "bbh": "BASELINE.json configs[4]: GeneralizedHarmonic, synthetic superposed Kerr-Schild data "
           "of two holes (m = 0.5 each, separation 16) on the BinaryCompactObject domain (44 "
           "blocks: per object six spherical and six cube wedges, ten bulged frustums, ten outer "
           "(half-)wedges, opening angle 120 degrees, both objects excised; block layout of "
           "Inspiral.yaml:54-101 with CubeScale 1, one N and 8^L elements per block, static "
           "maps), DirichletAnalytic boundaries, AnalyticChristoffel gauge, exponential filter "
           "(Alpha 36, HalfPower 420, Inspiral.yaml:219-224), AB3, dt=2e-4; strong scaling: the "
           "domain is fixed, elements are cut in block / Z-curve order",
}


def workload_config(args, world, cpu=False):
    workload = args.workload
    ref = weak_refinement(world, args.refine)
    extra = {}
    if workload == "kerr-schild-shell":
        lr = shell_radial_level(args, world)
        n_el = 6 * 4 ** args.refine * 2 ** lr // world
        ref = [args.refine, args.refine, lr]
        extra = {"inner_boundary": args.inner_boundary, "outer_boundary": args.outer_boundary,
                 "filter": None if args.no_filter else
                 {"Alpha": FILTER[0], "HalfPower": FILTER[1]}}
    elif workload == "bbh":
        n_el = -(-44 * 8 ** args.refine // world)
        ref = [args.refine] * 3
        extra = {"blocks": 44, "elements_total": 44 * 8 ** args.refine,
                 "filter": None if args.no_filter else {"Alpha": 36.0, "HalfPower": 420}}
    else:
        n_el = (2 ** args.refine) ** 3
    gauge = ("AnalyticChristoffel" if workload.startswith("kerr-schild") or workload == "bbh"
             else args.gauge)
    state_gb = n_el * 50 * args.points ** 3 * 8 / 1e9
    cfg = {
        "workload": WORKLOAD_NAMES[workload], **extra,
        "elements_per_gpu": n_el, "refinement": ref,
        "points_per_dim": args.points, "gauge": gauge, "stepper": "AdamsBashforth(3)",
        "parallelism": f"elements partitioned along the block Z-curve / by radius over {world} "
                       "GPU(s), mortar-face halo exchange (NCCL send/recv)",
        "cache": f"inputs larger than L2 (state {state_gb:.2f} GB + 3 history slots per GPU; no "
                 "flush needed)" if not cpu else "n/a (CPU arm)",
    }
    if cpu:
        sr = args.cpu_sample_refine
        cfg["cpu_sample"] = (f"bounded sample of this workload: refinement "
                             f"{(sr, sr + 1) if workload == 'kerr-schild-shell' else sr} "
                             "instead of the full element count (same N, physics, stepper, filter)")
    return cfg


def make_problem(evolution, args, world):
    N = args.points
    if args.workload == "kerr-schild-shell":
        return shell_problem(evolution, args.refine, shell_radial_level(args, world), N,
                             args.inner_boundary, args.outer_boundary)
    if args.workload == "bbh":
        return evolution.gh_binary_problem(args.refine, N)
    refinement = weak_refinement(world, args.refine)
    if args.workload == "kerr-schild":
        # element size fixed (1/8 M per element edge), lattice grows with the GPU count
        ne = [2 ** r for r in refinement]
        return evolution.gh_kerr_schild_problem(
            refinement, N, lower=(2.0, 2.0, 2.0), upper=tuple(2.0 + 0.125 * n for n in ne))
    # weak scaling keeps the element size of the single-GPU run (1/2^refine): the
    # periodic domain grows by whole wavelengths, so dt stays inside the AB3 limit
    upper = tuple(float(2 ** (r - args.refine)) for r in refinement)
    return evolution.gh_gauge_wave_problem(refinement, N, upper=upper)


class GpuRun:
    """One workload on this rank's GPU: set-up, timed steps, per-kernel roofline."""

    def __init__(self, args, world, rank, local_rank, pg):
        import torch
        from spectre_b200 import evolution, lib
        self.torch, self.lib, self.evolution = torch, lib, evolution
        self.args, self.world, self.rank, self.local_rank, self.pg = args, world, rank, local_rank, pg
        self.problem = make_problem(evolution, args, world)
        self.gauge = (lib.GAUGE_HARMONIC if args.gauge == "harmonic"
                      else lib.GAUGE_ANALYTIC_GAUGE_WAVE)
        self.gauge_params = (0.1, 1.0) if args.gauge == "analytic" else ()
        self.use_filter = args.workload in ("kerr-schild-shell", "bbh") and not args.no_filter
        self.filter = (36.0, 420) if args.workload == "bbh" else FILTER
        self.ev = self.new_evolution()
        self.ctx = self.ev.ctx
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=f"cuda:{local_rank}")

    def new_evolution(self):
        lib = self.lib
        ev = self.evolution.Evolution(self.problem, lib.STEPPER_ADAMS_BASHFORTH, 3, self.args.dt,
                                      0.0, self.gauge, self.gauge_params, self.local_rank,
                                      self.world, self.rank, self.pg,
                                      native_exchange=not getattr(self.args, "python_exchange",
                                                                  False))
        if self.use_filter:
            ev.ctx.set_exponential_filter(True, *self.filter)
        if self.args.volume_variant:
            ev.ctx.set_split_volume(self.args.volume_variant)
        return ev

    def barrier(self):
        self.ctx.synchronize()
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            self.torch.cuda.synchronize()

    def timed_steps(self, steps, warmup, sample_clocks):
        torch, lib, ev = self.torch, self.lib, self.ev
        # self-start (not timed: SURVEY 8d "exclude init, self-start") + warm-up
        ev.take_steps(warmup)
        self.barrier()
        if os.environ.get("BENCH_NO_CLOCK_SAMPLER") == "1":
            sample_clocks = False
        sampler = ClockSampler(self.local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        launches0 = lib.kernel_launch_count()
        evals0 = self.ctx.rhs_evaluations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        ev.take_steps(steps)
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = lib.kernel_launch_count() - launches0
        assert self.ctx.rhs_evaluations - evals0 == steps
        self.phases, self.extra_steps = None, 0
        if self.world > 1 and getattr(ev, "native_exchange", False):
            # event timeline of one more (untimed) step of the multi-GPU schedule
            self.ctx.set_phase_timing(True)
            ev.take_steps(1)
            self.phases = self.ctx.phase_times()
            self.ctx.set_phase_timing(False)
            self.extra_steps = 1
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=f"cuda:{self.local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    def check_state(self):
        """The state must still be finite and close to the exact solution."""
        state = self.ctx.get_state()
        assert np.isfinite(state).all(), "state is not finite after the timed run"
        ids = self.ev.part.global_ids
        sel = np.unique(np.linspace(0, len(ids) - 1, 16).astype(int))
        exact = self.problem.u0(ids[sel], self.ctx.time)
        err = float(np.max(np.abs(state[sel] - exact)))
        if self.args.workload == "bbh":
            # superposed Kerr-Schild data are not a solution: the evolution moves away from
            # them (by design); only a gross failure is an error here
            assert err < 0.1, f"state moved implausibly far from the initial data: {err}"
        else:
            assert err < 1e-3, f"solution drifted from the exact solution: {err}"
        return state, err

    def roofline(self, value):
        args, N = self.args, self.args.points
        kms = self.ctx.time_kernels(reps=5, update_terms=3)
        peak, peak_src = peaks()
        # static per-point fields read by the volume kernel: 3 damping fields, +20
        # when the gauge source function comes from memory (G = 23)
        G = 23 if (args.workload.startswith("kerr-schild") or args.workload == "bbh"
                   or args.gauge == "analytic") else 3
        kb = kernel_alg_bytes(N, n_static=G)
        kb["volume_update_fused"] = kb["volume"] + kb["update"]
        names = ["face", "volume", "update", "volume_update_fused", "filter"]
        # the step launches the face kernel, the fused volume+update kernel and (if
        # enabled) the filter; the separate volume / update timings are for comparison
        in_step = ["face", "volume_update_fused"] + (["filter"] if self.use_filter else [])
        kms_d = {n: float(m) for n, m in zip(names, kms)}
        dom = max(in_step, key=lambda n: kms_d[n])
        pts = self.ev.n_points
        achieved = kb[dom] * pts / (kms_d[dom] * 1e-3) / 1e9
        kname = {"face": "gh_face_kernel", "filter": "exponential_filter_kernel",
                 "volume_update_fused": "gh_volume_kernel (stepper update fused)"}[dom]
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(f"{kname}|{N}|{self.ev.part.n_local}")
        except Exception:
            traffic = None
        step_time = sum(kms_d[n] for n in in_step)
        b = b_alg(N, G=G)
        step = {"b_alg_bytes_per_update": b, "achieved": value / self.world * b / 1e9,
                "frac": value / self.world * b / 1e9 / peak,
                "note": "whole step per GPU against SURVEY.md 8(d) B_alg"}
        if self.use_filter:
            bf = b + kb["filter"]
            step["note"] += ("; the step also runs the exponential filter pass (read u, write "
                             "u = 800 B/update) that B_alg does not count: frac_incl_filter_bytes "
                             "adds those bytes")
            step["frac_incl_filter_bytes"] = value / self.world * bf / 1e9 / peak
        return {
            "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "alg_bytes_per_launch": kb[dom] * pts,
            "alg_bytes_note": "SURVEY 8(d) accounting: volume group + update group (the fused "
                              "kernel actually moves less: u and dt_u are not re-read)",
            "kernels_ms": kms_d,
            "kernels_share_of_step": {n: kms_d[n] / step_time for n in in_step},
            "kernels_frac": {n: (kb[n] * pts / (kms_d[n] * 1e-3) / 1e9 / peak
                                 if kms_d[n] > 0 else None) for n in names},
            "step": step,
        }

    def e2e(self, state, total_points):
        """End to end through the C-ABI with host buffers: state H2D + one step + D2H."""
        import ctypes
        torch, lib, args, ev, ctx = self.torch, self.lib, self.args, self.ev, self.ctx
        world = self.world
        nbytes = state.nbytes
        host = torch.empty(state.size, dtype=torch.float64, pin_memory=True)
        host_np = host.numpy().reshape(state.shape)
        host_np[...] = state
        L = lib.load()
        k_e2e = max(3, min(args.steps, 5))
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            lib._check(L.dgrhs_set_state(ctx._h, ctypes.c_void_p(host.data_ptr())))
            ev.take_steps(1)
            lib._check(L.dgrhs_get_state(ctx._h, ctypes.c_void_p(host.data_ptr())))
        self.barrier()
        el = time.perf_counter() - t0
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([el], device=f"cuda:{self.local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = {"value": total_points * k_e2e / el, "unit": UNIT,
               "h2d_bytes_per_step": int(nbytes) * world, "d2h_bytes_per_step": int(nbytes) * world,
               "steps": k_e2e,
               "what": "dgrhs_set_state(pinned host) + one AB3 step + dgrhs_get_state per step"}
        if world == 1 and not args.no_e2e_pipeline:
            # the same three calls per batch in their stream-ordered form on several contexts
            # in rotation: one batch uploads while another steps or downloads (PCIe is full duplex);
            # every batch still crosses the bus both ways inside the timed region
            serial = e2e["value"]
            n_lanes = max(2, args.e2e_lanes)
            lanes = [(ev, host_np)]
            for _ in range(n_lanes - 1):
                ev_x = self.new_evolution()
                ev_x.take_steps(args.warmup)          # self-start outside the timed region
                host_x = torch.empty(state.size, dtype=torch.float64, pin_memory=True)
                host_x_np = host_x.numpy().reshape(state.shape)
                host_x_np[...] = state
                lanes.append((ev_x, host_x_np))
                self._pinned = getattr(self, "_pinned", []) + [host_x]
            k_pipe = n_lanes * k_e2e
            for e, _ in lanes:
                e.ctx.synchronize()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(k_pipe):
                e, h = lanes[i % n_lanes]
                e.ctx.synchronize()               # this lane's previous batch is back on the host
                e.ctx.set_state_async(h)
                e.take_steps(1)
                e.ctx.get_state_async(h)
            for e, _ in lanes:
                e.ctx.synchronize()
            el = time.perf_counter() - t0
            assert all(np.isfinite(h).all() for _, h in lanes)
            e2e.update({"value": total_points * k_pipe / el, "steps": k_pipe,
                        "serial_value": serial, "lanes": n_lanes,
                        "what": "per batch: dgrhs_set_state_async(pinned host) + one AB3 step + "
                                "dgrhs_get_state_async, %d contexts in rotation so that one "
                                "batch uploads while another steps and a third downloads (PCIe "
                                "is full duplex); every batch crosses the bus both ways inside "
                                "the timed region; serial_value = one context, blocking calls"
                                % n_lanes})
            for e, _ in lanes[1:]:
                e.ctx.close()
        return e2e

    def verify(self, steps_total):
        """Multi-GPU: gathered state bit-for-bit against a single-GPU evolution."""
        lib = self.lib
        n_global = self.problem.brick.n_elements
        gathered = self.ev.gather_state(n_global)
        if self.rank == 0:
            ref_ev = self.evolution.Evolution(self.problem, lib.STEPPER_ADAMS_BASHFORTH, 3,
                                              self.args.dt, 0.0, self.gauge, self.gauge_params,
                                              self.local_rank)
            if self.use_filter:
                ref_ev.ctx.set_exponential_filter(True, *self.filter)
            ref_ev.take_steps(steps_total)
            same = np.array_equal(ref_ev.gather_state(n_global), gathered)
            print(f"[verify] {self.world}-rank state bit-identical to single-GPU state: {same}",
                  file=sys.stderr)
            assert same, "multi-GPU evolution differs from the single-GPU evolution"
            ref_ev.ctx.close()


def secondary_line(args, world, rank, local_rank, pg):
    """BASELINE.json configs[1] in the same run (short): value, step time, roofline."""
    sub = argparse.Namespace(**vars(args))
    sub.workload = "gauge-wave"
    sub.refine, sub.points, sub.cpu_sample_refine = WORKLOAD_DEFAULTS["gauge-wave"]
    sub.gauge, sub.scaling = "harmonic", "weak"
    run = GpuRun(sub, world, rank, local_rank, pg)
    steps = 50
    ms, launches, _ = run.timed_steps(steps, 3, sample_clocks=False)
    total_points = run.ev.n_points * world
    value = total_points * steps / (ms * 1e-3)
    _, err = run.check_state()
    # at two ranks: the gathered state against a single-GPU evolution of the whole domain, bit
    # for bit (cheap on this workload; at more ranks the whole domain is not run on one GPU)
    bit_identical = None
    if world == 2:
        try:
            run.verify(3 + steps + run.extra_steps)
            bit_identical = True
        except AssertionError:
            bit_identical = False
    roof = run.roofline(value)
    run.ctx.close()
    return {"config": workload_config(sub, world), "value": value, "unit": UNIT, "steps": steps,
            "ms_per_step": ms / steps, "gpu_launches": int(launches),
            "two_rank_state_bit_identical_to_one_gpu": bit_identical,
            "roofline": {k: roof[k] for k in ("kernel", "achieved", "peak", "frac", "kernels_ms",
                                              "step")},
            "max_abs_error_vs_exact": err}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kerr-schild-shell",
                    choices=["gauge-wave", "kerr-schild", "kerr-schild-shell", "bbh"],
                    help="kerr-schild-shell: BASELINE configs[3], the north-star configuration "
                         "(default); gauge-wave: configs[1]; kerr-schild: Brick-lattice stand-in")
    ap.add_argument("--refine", type=int, default=None,
                    help="shell: 4^refine angular x 2^(refine+1) radial elements per wedge and "
                         "GPU (default 3: 6144 elements); bricks: 2^refine elements per dim "
                         "per GPU (default 4)")
    ap.add_argument("--points", type=int, default=None,
                    help="LGL points per dimension N = P+1 (default: 12 shell, 8 gauge wave)")
    ap.add_argument("--dt", type=float, default=2e-4)
    ap.add_argument("--gauge", default="harmonic", choices=["harmonic", "analytic"],
                    help="gauge-wave workload only (the Kerr-Schild workloads use "
                         "AnalyticChristoffel)")
    ap.add_argument("--no-filter", action="store_true",
                    help="kerr-schild-shell: switch the exponential filter off")
    ap.add_argument("--cpu-sample-refine", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lts", action="store_true",
                    help="skip the local-time-stepping side measurement (N = 1 only)")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the configs[1] measurement reported as `secondary`")
    ap.add_argument("--e2e-lanes", type=int, default=3,
                    help="contexts in rotation in the pipelined e2e measurement (N = 1)")
    ap.add_argument("--no-e2e-pipeline", action="store_true",
                    help="measure e2e with blocking calls on one context only")
    ap.add_argument("--outer-boundary", default="DirichletAnalytic",
                    choices=["DirichletAnalytic", "ConstraintPreserving",
                             "ConstraintPreservingPhysical"],
                    help="kerr-schild-shell: boundary condition on the outer sphere")
    ap.add_argument("--inner-boundary", default="DirichletAnalytic",
                    choices=["DirichletAnalytic", "DemandOutgoingCharSpeeds"],
                    help="kerr-schild-shell: boundary condition on the excision sphere")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="kerr-schild-shell: weak = 2^(refine+1) radial elements per GPU; "
                         "strong = 2^(refine+1) * strong-factor radial elements in total")
    ap.add_argument("--strong-factor", type=int, default=4)
    ap.add_argument("--volume-variant", type=int, default=0,
                    help="dgrhs_set_split_volume: 0 default, 1 split kernels (A/B comparisons)")
    ap.add_argument("--python-exchange", action="store_true",
                    help="multi-GPU A/B: drive the halo exchange from Python over "
                         "torch.distributed (round-1 schedule) instead of inside libdgrhs.so")
    ap.add_argument("--verify", action="store_true",
                    help="multi-GPU: compare the gathered state bit-for-bit with a single-GPU "
                         "evolution of the same global problem on rank 0 (small configs)")
    args = ap.parse_args()
    d_refine, d_points, d_sample = WORKLOAD_DEFAULTS[args.workload]
    args.refine = d_refine if args.refine is None else args.refine
    args.points = d_points if args.points is None else args.points
    args.cpu_sample_refine = d_sample if args.cpu_sample_refine is None else args.cpu_sample_refine
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference "
                         "for the CPU arm")
    torch.cuda.set_device(local_rank)
    pg = None
    if world > 1:
        # one process per GPU: run on the CPUs next to that GPU, so that the pinned host
        # buffers of the e2e leg are first touched on the GPU's NUMA node (the ranks otherwise
        # share one node's memory controllers); the CPU baseline only runs at N = 1
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:   # no NVML / not permitted: keep the inherited affinity
            pass
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        pg = dist.group.WORLD

    run = GpuRun(args, world, rank, local_rank, pg)
    ms, launches, clocks = run.timed_steps(args.steps, args.warmup, sample_clocks=rank == 0)
    total_points = run.ev.n_points * world
    value = total_points * args.steps / (ms * 1e-3)
    state, err = run.check_state()
    if args.verify and world > 1:
        run.verify(args.warmup + args.steps + run.extra_steps)
    roofline = run.roofline(value)
    e2e = None if args.no_e2e else run.e2e(state, total_points)
    del state
    run.ctx.close()

    secondary = None
    if args.workload == "kerr-schild-shell" and not args.no_secondary:
        secondary = secondary_line(args, world, rank, local_rank, pg)

    # local time stepping (not the metric: a side measurement of the same library, N = 1 only)
    lts = None
    if rank == 0 and world == 1 and args.workload == "kerr-schild-shell" and not args.no_lts:
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "r02_lts_shell", os.path.join(ROOT, "profiles", "r02_lts_shell.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        try:
            lts = mod.measure()
            lts["what"] = ("Adams-Bashforth local time stepping (dgrhs_lts_*, DESIGN.md 3.6) "
                           "against GTS with the finest step, wall time for the same simulated "
                           "time on a shell whose radial element size grows 16x")
        except Exception as e:   # the headline does not depend on this leg
            lts = {"error": str(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(args.workload, args.points, args.cpu_sample_refine, args.dt, run.use_filter)
        cpu = arm.run(40, 1, budget_s=20.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "value_per_core")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": ("strong" if args.workload == "bbh" else
                        args.scaling if args.workload == "kerr-schild-shell" else "weak"),
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "max_abs_error_vs_exact": err, "secondary": secondary, "lts": lts,
        }
        if run.phases:
            line["multi_gpu_timeline_ms"] = {
                "rank": 0, "what": "CUDA-event timeline of one RHS evaluation inside "
                "dgrhs_take_steps, ms since its start (main stream: pack, faces_interior, "
                "volume_interior; comm stream: nccl_start, nccl_end, faces_boundary, "
                "volume_boundary)", **run.phases}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
