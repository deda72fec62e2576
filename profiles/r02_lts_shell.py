"""Round 2: local time stepping at scale (not the bench.py metric; a side measurement).
GH Kerr-Schild on a thick shell (inner radius 1.9 M, outer radius 30.4 M, Logarithmic radial
distribution: the radial element size grows 16x from the inside out), AB3.  GTS has to take the
step of the innermost layer everywhere; LTS gives each radial layer the largest power-of-two
multiple of it that StepChoosers::ElementSizeCfl allows (evaluated once, at the start).  Both runs go through libdgrhs.so; the
LTS kernels (snapshot / boundary / add) are not tuned.  Writes one JSON line.

    python profiles/r02_lts_shell.py [--points 10] [--angular 3] [--radial 4]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectre_b200 import evolution, lib  # noqa: E402
from spectre_b200 import lts as hlts  # noqa: E402


def measure(points=10, angular=3, radial=4, coarse_steps=2, dt_fine=2e-4):
    """GTS (finest step everywhere) and LTS wall time for the same simulated time; returns the
    dict that main() prints (bench.py reports it as its `lts` object at N = 1)."""
    args = argparse.Namespace(points=points, angular=angular, radial=radial,
                              coarse_steps=coarse_steps, dt_fine=dt_fine)
    N, order = args.points, 3
    problem = evolution.gh_kerr_schild_shell_problem((args.angular, args.radial), N,
                                                     inner_radius=1.9, outer_radius=30.4,
                                                     order="radial")
    ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, order, args.dt_fine)
    part = ev.part
    ids = part.global_ids
    x, J, stat = problem.coords(ids), problem.inverse_jacobian(ids), problem.static(ids)
    u0 = problem.u0(ids, 0.0)
    nelem = len(ids)
    # StepChoosers::ElementSizeCfl evaluated at the start gives the ratios of the element steps
    # (spectre_b200/lts.py); the finest step is --dt-fine
    speed = hlts.gh_largest_characteristic_speed(u0, stat[:, 1])
    goal = hlts.element_size_cfl(hlts.size_of_element(problem.brick, ids), speed,
                                 lib.stepper_properties(lib.STEPPER_ADAMS_BASHFORTH, order)[3],
                                 1.0)
    fine = np.floor(np.log2(goal / goal.min() * (1 + 1e-9))).astype(int)   # log2 of step / finest
    lmax = int(fine.max())
    levels = (lmax - fine).astype(np.int32)
    perm, nb = hlts.order_by_level(levels, part.local_neighbors)
    levels = levels[perm]
    dt_coarse = args.dt_fine * 2 ** lmax
    ghost = ev.boundary_ghost_data(problem, 0.0)

    # ---- GTS with the fine step (the tuned path), same number of simulated time units
    ctx = ev.ctx
    ctx.set_state(u0)
    n_fine = args.coarse_steps * 2 ** lmax
    ctx.take_steps(8)
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.take_steps(n_fine)
    ctx.synchronize()
    gts_s = time.perf_counter() - t0
    err_gts = float(np.max(np.abs(ctx.get_state() - u0)))
    ctx.close()

    # ---- LTS on the reordered elements
    ctx = lib.Context(lib.SYSTEM_GH, N, nelem, len(part.external_faces))
    ctx.set_geometry(J[perm], x[perm], nb)
    ctx.set_neighbor_orientations(part.local_neighbor_direction[perm],
                                  part.local_face_permutation[perm])
    ctx.set_static_fields(stat[perm])
    ctx.set_gauge(lib.GAUGE_FIELDS)
    ctx.set_gauge_analytic_christoffel(u0[perm])
    ctx.set_boundary_ghost_data(0, ghost)
    ctx.set_state(u0[perm])
    ctx.lts_init(order, 0.0, dt_coarse, levels)
    for j in range(1, order):
        ctx.lts_set_past_state(j, u0[perm])      # static solution
    ctx.lts_take_coarse_steps(1)
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.lts_take_coarse_steps(args.coarse_steps)
    ctx.synchronize()
    lts_s = time.perf_counter() - t0
    err_lts = float(np.max(np.abs(ctx.get_state() - u0[perm])))
    ctx.close()

    counts = {int(l): int((levels == l).sum()) for l in sorted(set(levels.tolist()))}
    updates_lts = sum(c * 2 ** l for l, c in counts.items()) * N ** 3 * args.coarse_steps
    updates_gts = nelem * 2 ** lmax * N ** 3 * args.coarse_steps
    return {
        "workload": "GH Kerr-Schild, shell 1.9 M .. 30.4 M, Logarithmic, AB3, N=%d, %d elements"
                    % (N, nelem),
        "elements_per_level": counts, "dt_fine": args.dt_fine, "dt_coarse": dt_coarse,
        "simulated_time": args.coarse_steps * dt_coarse,
        "gts_seconds": gts_s, "lts_seconds": lts_s, "lts_speedup_wall": gts_s / lts_s,
        "element_updates_gts": updates_gts, "element_updates_lts": updates_lts,
        "work_ratio": updates_gts / updates_lts,
        "gts_updates_per_s": updates_gts / gts_s, "lts_updates_per_s": updates_lts / lts_s,
        "max_drift_from_static_solution": {"gts": err_gts, "lts": err_lts}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10)
    ap.add_argument("--angular", type=int, default=3)
    ap.add_argument("--radial", type=int, default=4)
    ap.add_argument("--coarse-steps", type=int, default=2)
    ap.add_argument("--dt-fine", type=float, default=2e-4)
    a = ap.parse_args()
    print(json.dumps(measure(a.points, a.angular, a.radial, a.coarse_steps, a.dt_fine)))


if __name__ == "__main__":
    main()
