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


def levels_from_step_limit(step_limit, dt_coarse, max_level=7):
    """Smallest level whose step dt_coarse / 2^level does not exceed the element's step limit
    (the role of the reference's StepChoosers::ElementSizeCfl / Cfl, which bound the step by
    the element size over the characteristic speed; here evaluated once, at the start)."""
    lim = np.asarray(step_limit, dtype=float)
    lv = np.ceil(np.log2(np.maximum(dt_coarse / lim, 1.0)) - 1e-12).astype(int)
    return np.clip(lv, 0, max_level)


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
