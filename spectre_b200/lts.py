"""Host-side helpers of the local-time-stepping entry points (dgrhs_lts_*): the library
wants the elements sorted by step-size level (coarse steps first), so that the elements with
a step boundary at a given time form a suffix of the element order."""
from __future__ import annotations

import numpy as np


def order_by_level(levels, neighbors):
    """(perm, neighbors'): element order sorted by level (stable) and the neighbour table
    [n_elements, 6] renumbered for it (negative entries -- external faces, ghost slots --
    are kept).  Per-element arrays go along as a[perm]."""
    levels = np.asarray(levels)
    perm = np.argsort(levels, kind="stable")
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    nb = np.asarray(neighbors)[perm].copy()
    m = nb >= 0
    nb[m] = inv[nb[m]]
    return perm, nb.astype(np.int32)


def choose_lts_step_power(desired_step, slab_duration, fraction_denominator=1):
    """choose_lts_step_size (Time/ChooseLtsStepSize.cpp:14-39) for forward steps: n such that
    the step is slab / 2^n -- the largest binary fraction of the slab that does not exceed the
    desired step, but no larger than the position inside the slab (a time whose slab fraction
    has the denominator 2^m) allows.  Works on arrays."""
    count = slab_duration / np.asarray(desired_step, dtype=float)
    power = np.where(count == 0.0, 0.0, np.ceil(np.log2(np.maximum(np.ceil(count), 1.0))))
    steps = np.maximum(2.0 ** power, float(fraction_denominator))
    return np.log2(steps).astype(int)


def levels_from_step_limit(step_limit, dt_coarse, max_level=7):
    """Step-size level of every element for a per-element step limit (the goal of
    StepChoosers::ElementSizeCfl / Cfl, evaluated once, at the start): the element takes the
    steps dt_coarse / 2^level that choose_lts_step_size picks at the start of a slab of length
    dt_coarse."""
    return np.clip(choose_lts_step_power(step_limit, dt_coarse), 0, max_level)


def start_from_gts(ctx, order, t0, dt_coarse, levels, stepper=None):
    """Start an LTS evolution without analytic past states: the context (holding u(t0)) takes
    (order - 1) coarse steps' worth of global steps with the finest step (self-started
    Adams-Bashforth of the same order, the library's GTS path), the states at the times
    T_s - j * (element step) are collected per level and handed to dgrhs_lts_set_past_state;
    the LTS evolution then starts at T_s = t0 + (order - 1) dt_coarse.  (The reference
    self-starts all elements with one common step and lets the step choosers spread the step
    sizes afterwards; with fixed levels the common-step phase is this GTS phase.)
    Returns T_s."""
    from . import lib
    levels = np.asarray(levels)
    lmax = int(levels.max())
    stride = 2 ** (lmax - levels)
    tick = dt_coarse / 2 ** lmax
    n_ticks = (order - 1) * 2 ** lmax
    needed = {n_ticks - j * int(s) for s in set(stride.tolist()) for j in range(1, order)}
    ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH if stepper is None else stepper, order, t0, tick)
    snaps = {0: ctx.get_state()} if 0 in needed else {}
    for T in range(1, n_ticks + 1):
        ctx.take_steps(1)
        if T in needed:
            snaps[T] = ctx.get_state()
    t_start = t0 + n_ticks * tick
    ctx.lts_init(order, t_start, dt_coarse, levels)
    for j in range(1, order):
        past = np.empty_like(next(iter(snaps.values())))
        for s in set(stride.tolist()):
            sel = stride == s
            past[sel] = snaps[n_ticks - j * int(s)][sel]
        ctx.lts_set_past_state(j, past)
    return t_start


# ---- step choosers evaluated once, at the start -------------------------------------------
def size_of_element(dom, ids=None):
    """domain::size_of_element (Domain/SizeOfElement.cpp:44-59): per element and logical
    dimension the distance between the inertial centres of the two opposite faces.  Uses the
    domain's exact map (`map_points`) when it has one, else the affine bounds of a Brick."""
    ids = list(range(dom.n_elements)) if ids is None else list(ids)
    out = np.zeros((len(ids), 3))
    for k, e in enumerate(ids):
        for d in range(3):
            if hasattr(dom, "map_points"):
                xi = np.zeros((3, 2))
                xi[d] = (-1.0, 1.0)
                x, _ = dom.map_points(e, xi)
                out[k, d] = np.sqrt(((x[:, 1] - x[:, 0]) ** 2).sum())
            else:
                x = dom.coords([e])[0]
                out[k, d] = x[d].max() - x[d].min()     # affine: face-to-face distance
    return out


def gh_largest_characteristic_speed(u, gamma1):
    """gh::Tags::ComputeLargestCharacteristicSpeed (GeneralizedHarmonic/Characteristics.cpp:
    188-196) per element: max(|1 + gamma1| |beta|, |beta| + lapse) over the grid points, with
    lapse and shift from the spacetime metric.  u [nelem, 50, n], gamma1 [nelem, n]."""
    sym = {}
    s = 0
    for a in range(4):
        for b in range(a, 4):
            sym[(a, b)] = sym[(b, a)] = s
            s += 1
    g = lambda a, b: u[:, sym[(a, b)]]
    gam = np.stack([np.stack([g(i + 1, j + 1) for j in range(3)], axis=-1) for i in range(3)],
                   axis=-2)                                    # [nelem, n, 3, 3]
    beta_lo = np.stack([g(0, i + 1) for i in range(3)], axis=-1)
    beta_up = np.linalg.solve(gam, beta_lo[..., None])[..., 0]
    b2 = (beta_up * beta_lo).sum(axis=-1)
    lapse = np.sqrt(b2 - g(0, 0))
    mag = np.sqrt(b2)
    return np.maximum((np.abs(1.0 + gamma1) * mag).max(axis=1), (mag + lapse).max(axis=1))


def element_size_cfl(element_sizes, speed, stable_step, safety_factor):
    """StepChoosers::ElementSizeCfl (Time/StepChoosers/ElementSizeCfl.hpp:76-92): the step
    goal safety_factor * stable_step * min_d(size_d) / (speed * 3) of every element."""
    return safety_factor * stable_step * np.min(element_sizes, axis=1) / (np.asarray(speed) * 3)


class LtsEvolution:
    """Adams-Bashforth local time stepping of a `evolution.Problem` on one GPU: the elements are
    put into the order of their step-size levels, the library's context is set up like the GTS
    `evolution.Evolution` (geometry, orientations, mortars, boundary conditions, gauge), the
    histories start from the problem's analytic data at the elements' own past step times
    (past="analytic") or from a GTS phase with the finest step (past="gts").

    step_goal: per-element upper bound of the step (array in the problem's element order), or
    None for StepChoosers::ElementSizeCfl with `safety_factor`, evaluated once at t0; every
    element takes the largest step slab / 2^n that does not exceed min(step_goal, max_step)."""

    def __init__(self, problem, order, slab, t0=0.0, step_goal=None, safety_factor=0.5,
                 max_step=None, past="analytic", gauge=None, gauge_params=(), device=0,
                 filter_params=None):
        from . import evolution, lib
        self.problem, self.order, self.t0 = problem, int(order), float(t0)
        box = {}

        def order_elements(part, prob):
            ids = part.global_ids
            if step_goal is None:
                if prob.system != lib.SYSTEM_GH:
                    speed = np.ones(len(ids))       # ScalarWave: unit characteristic speed
                else:
                    speed = gh_largest_characteristic_speed(prob.u0(ids, t0),
                                                            prob.static(ids)[:, 1])
                stable = lib.stepper_properties(lib.STEPPER_ADAMS_BASHFORTH, self.order)[3]
                goal = element_size_cfl(size_of_element(prob.brick, ids), speed, stable,
                                        safety_factor)
            else:
                goal = np.asarray(step_goal, dtype=float)[ids]
            if max_step is not None:
                goal = np.minimum(goal, max_step)
            n = levels_from_step_limit(goal, slab, max_level=60)     # step = slab / 2^n
            box["n_min"] = int(n.min())
            levels = n - box["n_min"]
            if levels.max() > 7:
                raise ValueError("more than eight step-size levels")
            perm = np.argsort(levels, kind="stable")
            box["levels"] = levels[perm].astype(np.int32)
            return perm
        kw = {} if gauge is None else {"gauge": gauge, "gauge_params": gauge_params}
        self.dt_coarse = None
        self.ev = evolution.Evolution(problem, lib.STEPPER_ADAMS_BASHFORTH, self.order,
                                      slab, t0, device=device, element_order=order_elements, **kw)
        self.levels = box["levels"]
        self.dt_coarse = slab / 2 ** box["n_min"]
        self.ctx, self.part = self.ev.ctx, self.ev.part
        if filter_params:
            self.ctx.set_exponential_filter(True, *filter_params)
        ids = self.part.global_ids
        if past == "gts":
            self.ctx.set_stepper(lib.STEPPER_ADAMS_BASHFORTH, self.order, t0,
                                 self.dt_coarse / 2 ** int(self.levels.max()))
            self.t_start = start_from_gts(self.ctx, self.order, t0, self.dt_coarse, self.levels)
        else:
            self.ctx.lts_init(self.order, t0, self.dt_coarse, self.levels)
            for j in range(1, self.order):
                u = np.empty((len(ids), self.ctx.n_vars, self.ctx.n))
                for lv in sorted(set(self.levels.tolist())):
                    sel = self.levels == lv
                    u[sel] = problem.u0(ids[sel], t0 - j * self.dt_coarse / 2 ** lv)
                self.ctx.lts_set_past_state(j, u)
            self.t_start = t0

    @property
    def time(self):
        return self.ctx.lts_time()[0]

    def take_coarse_steps(self, n):
        if self.part.external_faces and self.problem.boundary_time_dependent:
            # time-dependent ghost data: refresh at every tick of the finest step
            per = 2 ** int(self.levels.max())
            for _ in range(n * per):
                self.ctx.set_boundary_ghost_data(
                    self.part.n_recv, self.ev.boundary_ghost_data(self.problem, self.time))
                self.ctx.lts_take_ticks(1)
        else:
            self.ctx.lts_take_coarse_steps(n)

    def state(self):
        """(global element ids, state) in the library's element order"""
        return self.part.global_ids, self.ctx.get_state()


def minimum_grid_spacing(coords, N):
    """domain::minimum_grid_spacing (Domain/MinimumGridSpacing.cpp:31-71) per element: the
    smallest inertial distance between a grid point and one of its 26 index neighbours.
    coords [nelem, 3, N^3], xi fastest."""
    x = np.asarray(coords).reshape(len(coords), 3, N, N, N)      # [e, dim, k, j, i]
    best = np.full(len(coords), np.inf)
    for dk in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                if (dk, dj, di) <= (0, 0, 0):
                    continue        # the opposite offsets are seen from the other point
                a = x[:, :, max(0, -dk):N - max(0, dk), max(0, -dj):N - max(0, dj),
                      max(0, -di):N - max(0, di)]
                b = x[:, :, max(0, dk):N - max(0, -dk), max(0, dj):N - max(0, -dj),
                      max(0, di):N - max(0, -di)]
                d = np.sqrt(((a - b) ** 2).sum(axis=1)).reshape(len(coords), -1).min(axis=1)
                best = np.minimum(best, d)
    return best


def cfl_step(min_grid_spacing, speed, stable_step, safety_factor):
    """StepChoosers::Cfl (Time/StepChoosers/Cfl.hpp:69-79): safety_factor * stable_step *
    minimum grid spacing / (speed * 3)."""
    return safety_factor * stable_step * np.asarray(min_grid_spacing) / (np.asarray(speed) * 3)


def limit_increase(last_step, factor):
    """StepChoosers::LimitIncrease (LimitIncrease.hpp): the next step may be at most `factor`
    times the last one (a size limit, not a goal)."""
    return np.abs(last_step) * factor
