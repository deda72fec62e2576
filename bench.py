#!/usr/bin/env python
"""bench.py -- fp64 DG grid-point RHS updates/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # CPU arm (the oracle port)

Workload (config.workload): BASELINE.json configs[1] -- GeneralizedHarmonic
gauge wave on the periodic Brick [0,1]^3, 16^3 elements per GPU, N = P+1 = 8
Legendre-Gauss-Lobatto points per dimension, Adams-Bashforth 3, dt = 2e-4,
gamma0/1/2 = 1/-1/1, UpwindPenalty.  A "step" is one full AB3 time step of all
elements = one RHS evaluation per grid point + the stepper update.  Multi-GPU
runs are weak scaling (16^3 elements per rank, halo exchange over NCCL).

The JSON line carries: value (state resident in HBM), e2e (state copied
host->device and back every step through the C-ABI), roofline of the dominant
kernel (CUDA-event timed inside this run), cpu_baseline (oracle port on the
host cores, bounded sample), clocks, gpu_launches.
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


def b_alg(N, n_u=50, G=3, T_f=6, k=3):
    """Algorithmic bytes per grid-point update, SURVEY.md 8(d)."""
    return 8.0 * ((2 * n_u + 9 + G) + (k + 2) * n_u + (6.0 / N) * (2 * (n_u + T_f) + 2 * n_u))


def kernel_alg_bytes(N, n_u=50, n_static=3, k=3):
    """Compulsory bytes per grid point of each of OUR kernels:
    face: both sides' face values + 5 face statics in, lifted corrections out;
    volume: u, J^-1, statics, corrections in, dt_u out; update: u, k derivs in,
    u out."""
    face = 8.0 * (6.0 / N) * (2 * (n_u + 5) + n_u)
    volume = 8.0 * (2 * n_u + 9 + n_static + (6.0 / N) * n_u)
    update = 8.0 * (k + 2) * n_u
    return {"face": face, "volume": volume, "update": update}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def weak_refinement(world, base):
    """base^3 elements per rank: double one dimension per factor of two."""
    ref = [base, base, base]
    w, d = world, 0
    while w > 1:
        ref[d % 3] += 1
        w //= 2
        d += 1
    return ref


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample of the workload)
# ---------------------------------------------------------------------------
def cpu_oracle_run(N, sample_refine, steps, warmup, dt, budget_s=25.0):
    from oracle import oracle as orc
    orc.use_optimized_build()   # -O3 -march=native timing build (never used for parity)
    b = orc.Brick([0, 0, 0], [1, 1, 1], [sample_refine] * 3, N)
    x, J, nb = b.coords(), b.inverse_jacobian(), b.neighbors()
    u = np.stack([orc.gh_vars_from_metric(*orc.gauge_wave_metric(x[e], 0.0))
                  for e in range(b.nelem)])
    stat = np.zeros((b.nelem, 3, b.n))
    stat[:, 0], stat[:, 1], stat[:, 2] = 1.0, -1.0, 1.0
    ev = orc.Evolution(lambda v, t: orc.dg_rhs(1, N, v, J, stat, nb), u, 0.0, dt, "AB3")
    for _ in range(warmup):
        ev.step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        ev.step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    el = time.perf_counter() - t0
    pts = b.nelem * b.n
    return {"value": pts * done / el, "unit": UNIT, "cores": int(orc.lib().orc_num_threads()),
            "kind": "port",
            "sample": f"{b.nelem} elements (refinement {sample_refine}), N={N}, {done} AB3 steps "
                      f"in {el:.1f} s, oracle/dg_oracle.c (gcc -O3 -march=native) + numpy update, "
                      f"OpenMP over elements",
            "steps": done, "ms_per_step": 1e3 * el / done}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_oracle_run(args.points, args.cpu_sample_refine, args.steps, min(args.warmup, 1),
                       args.dt, budget_s=60.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": min(args.warmup, 1),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1, cpu=True),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "reference Charm++ executable cannot be built here (no Charm++/Blaze/...): "
                "this arm times the CPU restatement (oracle) of the same path",
    }
    print(json.dumps(line))


def workload_config(args, world, cpu=False):
    ref = weak_refinement(world, args.refine)
    names = {
        "gauge-wave": "BASELINE.json configs[1]: GeneralizedHarmonic gauge wave (A=0.1, "
                      "lambda=1), periodic Brick [0,1]^3 per GPU (the domain grows by whole "
                      "wavelengths with the GPU count), AB3, dt=2e-4, UpwindPenalty, "
                      "gamma0/1/2=1/-1/1",
        "kerr-schild": "BASELINE.json configs[2]/[3] stand-in: GeneralizedHarmonic Kerr-Schild "
                       "(M=1, a=0) on a Brick lattice (elements of edge M/8 from x=2M), "
                       "DirichletAnalytic boundaries, AnalyticChristoffel gauge, "
                       "GaussianPlusConstant damping (KerrSchild.yaml), AB3, dt=2e-4",
        "kerr-schild-shell": "BASELINE.json configs[2]/[3]: GeneralizedHarmonic Kerr-Schild "
                             "(M=1, a=0) on the Sphere domain with excision (six equiangular "
                             "wedges per layer, Logarithmic radial distribution, inner radius "
                             "1.9 M, h-refined: 4^L angular x 2^Lr radial elements per wedge, "
                             "the shell grows outwards with the GPU count and is cut at constant radius), "
                             "DirichletAnalytic "
                             "boundaries, AnalyticChristoffel gauge, GaussianPlusConstant "
                             "damping (KerrSchild.yaml), AB3, dt=2e-4",
    }
    workload = getattr(args, "workload", "gauge-wave")
    if workload == "kerr-schild-shell":
        lr = shell_radial_level(args, world)
        n_el = 6 * 4 ** args.refine * 2 ** lr // world
        ref = [args.refine, args.refine, lr]
    else:
        n_el = (2 ** args.refine) ** 3
    extra = {}
    if workload == "kerr-schild-shell":
        extra = {"inner_boundary": args.inner_boundary, "outer_boundary": args.outer_boundary}
    return {
        "workload": names[workload], **extra,
        "elements_per_gpu": n_el, "refinement": ref,
        "points_per_dim": args.points, "gauge": args.gauge, "stepper": "AdamsBashforth(3)",
        "parallelism": f"elements partitioned along the block Z-curve over {world} GPU(s), "
                       "mortar-face halo exchange (NCCL send/recv)",
        "cache": "inputs larger than L2 (state 0.84 GB + 3 history slots per GPU; no flush "
                 "needed)" if not cpu else "n/a (CPU arm)",
    }


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refine", type=int, default=4, help="2^refine elements per dim per GPU")
    ap.add_argument("--points", type=int, default=8, help="LGL points per dimension (N = P+1)")
    ap.add_argument("--dt", type=float, default=2e-4)
    ap.add_argument("--gauge", default="harmonic", choices=["harmonic", "analytic"])
    ap.add_argument("--workload", default="gauge-wave", choices=["gauge-wave", "kerr-schild", "kerr-schild-shell"],
                    help="gauge-wave: BASELINE configs[1] (default, the headline); kerr-schild: "
                         "configs[2]/[3] stand-in (Kerr-Schild on a Brick lattice with "
                         "DirichletAnalytic boundaries, AnalyticChristoffel gauge)")
    ap.add_argument("--cpu-sample-refine", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
                    help="dgrhs_set_split_volume: 0 default, 1 split kernels, 2 pair-staged "
                         "kernel for N >= 10 (A/B comparisons)")
    ap.add_argument("--verify", action="store_true",
                    help="multi-GPU: compare the gathered state bit-for-bit with a single-GPU "
                         "evolution of the same global problem on rank 0 (small configs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from spectre_b200 import evolution, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference "
                         "for the CPU arm")
    torch.cuda.set_device(local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        pg = dist.group.WORLD
    N = args.points
    refinement = weak_refinement(world, args.refine)
    if args.workload == "kerr-schild-shell":
        # radial element ratio q such that elements are a quarter as deep as they are
        # wide at every radius; the outer radius grows with the number of radial
        # elements (4.1 M on one GPU, 868 M on eight for --refine 3)
        lr = shell_radial_level(args, world)
        q = 1.0 + 0.125 * np.pi / 2 ** args.refine
        problem = evolution.gh_kerr_schild_shell_problem(
            (args.refine, lr), N, inner_radius=1.9, outer_radius=1.9 * q ** (2 ** lr),
            order="radial", inner_boundary=args.inner_boundary,
            outer_boundary=args.outer_boundary)
    elif args.workload == "kerr-schild":
        # element size fixed (1/8 M per element edge), lattice grows with the GPU count
        ne = [2 ** r for r in refinement]
        problem = evolution.gh_kerr_schild_problem(
            refinement, N, lower=(2.0, 2.0, 2.0), upper=tuple(2.0 + 0.125 * n for n in ne))
    else:
        # weak scaling keeps the element size of the single-GPU run (1/2^refine): the
        # periodic domain grows by whole wavelengths, so dt stays inside the AB3 limit
        upper = tuple(float(2 ** (r - args.refine)) for r in refinement)
        problem = evolution.gh_gauge_wave_problem(refinement, N, upper=upper)
    gauge = lib.GAUGE_HARMONIC if args.gauge == "harmonic" else lib.GAUGE_ANALYTIC_GAUGE_WAVE
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, args.dt, 0.0, gauge,
                             (0.1, 1.0) if args.gauge == "analytic" else (), local_rank, world,
                             rank, pg)
    ctx = ev.ctx
    if args.volume_variant:
        ctx.set_split_volume(args.volume_variant)
    stream = torch.cuda.ExternalStream(ctx.stream, device=f"cuda:{local_rank}")

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # self-start (not timed: SURVEY 8d "exclude init, self-start") + warm-up
    ev.take_steps(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.kernel_launch_count()
    evals0 = ctx.rhs_evaluations
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ev.take_steps(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.kernel_launch_count() - launches0
    rhs_evals = ctx.rhs_evaluations - evals0
    assert rhs_evals == args.steps
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_points = ev.n_points * world
    value = total_points * args.steps / (ms * 1e-3)

    # sanity: the state must still be finite and close to the exact solution
    state = ctx.get_state()
    assert np.isfinite(state).all(), "state is not finite after the timed run"
    exact = problem.u0(ev.part.global_ids[:8], ctx.time)
    err = float(np.max(np.abs(state[:8] - exact)))
    assert err < 1e-3, f"solution drifted from the exact solution: {err}"

    if args.verify and world > 1:
        n_global = problem.brick.n_elements
        steps_total = args.warmup + args.steps
        gathered = ev.gather_state(n_global)
        if rank == 0:
            ref_ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, args.dt, 0.0,
                                         gauge, (0.1, 1.0) if args.gauge == "analytic" else (),
                                         local_rank)
            ref_ev.take_steps(steps_total)
            same = np.array_equal(ref_ev.gather_state(n_global), gathered)
            print(f"[verify] {world}-rank state bit-identical to single-GPU state: {same}",
                  file=sys.stderr)
            assert same, "multi-GPU evolution differs from the single-GPU evolution"
            ref_ev.ctx.close()

    # per-kernel roofline (CUDA events inside the library, same stream)
    kms = ctx.time_kernels(reps=5, update_terms=3)
    peak, peak_src = peaks()
    # static per-point fields read by the volume kernel: 3 damping fields, +20
    # when the gauge source function comes from memory (G = 23)
    G = 23 if (args.workload.startswith("kerr-schild") or args.gauge == "analytic") else 3
    kb = kernel_alg_bytes(N, n_static=G)
    kb["volume_update_fused"] = kb["volume"] + kb["update"]
    names = ["face", "volume", "update", "volume_update_fused"]
    # the step launches the face kernel and the fused volume+update kernel; the
    # separate volume / update timings are reported for comparison only
    in_step = ["face", "volume_update_fused"]
    kms_d = {n: float(m) for n, m in zip(names, kms)}
    dom_name = max(in_step, key=lambda n: kms_d[n])
    pts_local = ev.n_points
    achieved = kb[dom_name] * pts_local / (kms_d[dom_name] * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        kname = {"face": "gh_face_kernel",
                 "volume_update_fused": "gh_volume_kernel (stepper update fused)"}[dom_name]
        traffic = tj.get(f"{kname}|{N}|{ev.part.n_local}")
    except Exception:
        traffic = None
    roofline = {
        "bound": "hbm",
        "kernel": {"face": "gh_face_kernel",
                   "volume_update_fused": "gh_volume_kernel (stepper update fused)"}[dom_name],
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "alg_bytes_per_launch": kb[dom_name] * pts_local,
        "alg_bytes_note": "SURVEY 8(d) accounting: volume group + update group (the fused "
                          "kernel actually moves less: u and dt_u are not re-read)",
        "kernels_ms": kms_d,
        "kernels_frac": {n: kb[n] * pts_local / (kms_d[n] * 1e-3) / 1e9 / peak for n in names},
        "step": {"b_alg_bytes_per_update": b_alg(N, G=G),
                 "achieved": value / world * b_alg(N, G=G) / 1e9,
                 "frac": value / world * b_alg(N, G=G) / 1e9 / peak,
                 "note": "whole step per GPU against SURVEY.md 8(d) B_alg"},
    }

    # end to end through the C-ABI with host buffers: state H2D + one step + D2H
    e2e = None
    if not args.no_e2e:
        nbytes = state.nbytes
        host = torch.empty(state.size, dtype=torch.float64, pin_memory=True)
        host_np = host.numpy().reshape(state.shape)
        host_np[...] = state
        import ctypes
        L = lib.load()
        k_e2e = max(3, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            lib._check(L.dgrhs_set_state(ctx._h, ctypes.c_void_p(host.data_ptr())))
            ev.take_steps(1)
            lib._check(L.dgrhs_get_state(ctx._h, ctypes.c_void_p(host.data_ptr())))
        barrier()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = {"value": total_points * k_e2e / el, "unit": UNIT,
               "h2d_bytes_per_step": int(nbytes) * world, "d2h_bytes_per_step": int(nbytes) * world,
               "steps": k_e2e,
               "what": "dgrhs_set_state(pinned host) + one AB3 step + dgrhs_get_state per step"}
        if world == 1 and not args.no_e2e_pipeline:
            # the same three calls per batch in their stream-ordered form on two contexts:
            # batch i+1 uploads while batch i steps and downloads (PCIe is full duplex);
            # every batch still crosses the bus both ways inside the timed region
            serial = e2e["value"]
            ev_b = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, 3, args.dt, 0.0, gauge,
                                       (0.1, 1.0) if args.gauge == "analytic" else (), local_rank,
                                       world, rank, pg)
            ev_b.take_steps(args.warmup)          # self-start outside the timed region
            host_b = torch.empty(state.size, dtype=torch.float64, pin_memory=True)
            host_b_np = host_b.numpy().reshape(state.shape)
            host_b_np[...] = state
            lanes = [(ev, host_np), (ev_b, host_b_np)]
            k_pipe = 2 * k_e2e
            for e, _ in lanes:
                e.ctx.synchronize()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(k_pipe):
                e, h = lanes[i % 2]
                e.ctx.synchronize()               # this lane's previous batch is back on the host
                e.ctx.set_state_async(h)
                e.take_steps(1)
                e.ctx.get_state_async(h)
            for e, _ in lanes:
                e.ctx.synchronize()
            el = time.perf_counter() - t0
            assert np.isfinite(host_np).all() and np.isfinite(host_b_np).all()
            e2e.update({"value": total_points * k_pipe / el, "steps": k_pipe,
                        "serial_value": serial,
                        "what": "per batch: dgrhs_set_state_async(pinned host) + one AB3 step + "
                                "dgrhs_get_state_async, two contexts double-buffered so one "
                                "batch uploads while the other steps and downloads; "
                                "serial_value = one context, blocking calls"})
            del ev_b

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_oracle_run(N, args.cpu_sample_refine, 40, 1, args.dt, budget_s=20.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling if args.workload == "kerr-schild-shell" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "max_abs_error_vs_exact": err,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
