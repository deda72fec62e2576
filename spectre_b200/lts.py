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
