"""TEST INFRASTRUCTURE (CPU oracle): local time stepping with Adams-Bashforth, restated from
the reference.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this.

* `lts_coefficients`  -- TimeSteppers::adams_lts::lts_coefficients for explicit schemes
  (src/Time/TimeSteppers/AdamsLts.cpp:165-437), in exact rational arithmetic.  Pinned to the
  known answers of tests/Unit/Time/TimeSteppers/Test_AdamsLts.cpp:416-760 (explicit cases).
* `LtsEvolution`      -- elements with step sizes slab / 2^level: the volume part of the time
  derivative (incl. external boundary conditions, which the reference applies inside
  ComputeTimeDerivative) goes through the element's own Adams-Bashforth history
  (Actions/UpdateU.hpp:44-120); the boundary corrections of internal faces are integrated
  over the step from the histories of both sides' packaged data
  (AdamsBashforth::add_boundary_delta_impl, AdamsBashforth.cpp:264-281;
  ApplyBoundaryCorrections.hpp:797-1010 with local_time_stepping == true: lifted_data = 0,
  add_boundary_delta, add_slice_to_data).  Step sizes are fixed per element (the state the
  step choosers of the reference reach when they stop changing the steps); the histories
  are initialised from given past states like TimeStepperTestUtils::initialize_history.
  Parity unpinned end to end (no reference executable here): pinned by composition -- the
  coefficients to the reference's known answers, equal levels to the GTS oracle, and the
  conservation / convergence properties the reference's own LTS tests use.
"""
from __future__ import annotations

from fractions import Fraction as Fr

import numpy as np

from . import oracle as orc


# --------------------------------------------------------------------------------------
# coefficients
# --------------------------------------------------------------------------------------
def _ab_exact(control, start, end):
    """integral over [start, end] of the Lagrange polynomials of `control`
    (adams_coefficients::variable_coefficients, AdamsCoefficients.cpp:73-112), exact"""
    out = []
    for j, tj in enumerate(control):
        poly = [Fr(1)]
        for m, tm in enumerate(control):
            if m == j:
                continue
            denom = tj - tm
            new = [Fr(0)] * (len(poly) + 1)
            for i, a in enumerate(poly):   # poly * (t - tm) / denom
                new[i + 1] += a / denom
                new[i] -= a * tm / denom
            poly = new
        val = Fr(0)
        for i, a in enumerate(poly):
            val += a * (end ** (i + 1) - start ** (i + 1)) / (i + 1)
        out.append(val)
    return out


def _lagrange_exact(control, x):
    """AdamsLts.cpp:280-305 (interpolation_coefficients)"""
    out = []
    for j, tj in enumerate(control):
        v = Fr(1)
        for m, tm in enumerate(control):
            if m != j:
                v *= (tm - x) / (tm - tj)
        out.append(v)
    return out


def _exact_time(entry):
    """adams_lts::exact_substep_time (AdamsLts.cpp:29-38): a step id (t, 0) is at t, the
    substep id (t, size) of the step that starts at t is at its end t + size"""
    return entry[0] + entry[1]


def _steps_of(entries):
    """the BoundaryHistory view of a flat list of ids: [(step entry, its substep entry or
    None)] in insertion order"""
    steps = []
    for en in entries:
        if en[1] == 0:
            steps.append([en, None])
        else:
            assert steps and steps[-1][0][0] == en[0], "substep without its step"
            steps[-1][1] = en
    return steps


def _relevant(entries, end, order, implicit=False):
    """AdamsLts.cpp:173-206 (find_relevant_ids): the most recent step ids before `end`, by
    position (the times need not be sorted during self-start) -- `order` of them for an
    explicit scheme, order - 1 plus the predictor value (substep) of the last one for an
    implicit scheme"""
    steps = _steps_of(entries)
    used_end = len(steps)
    while used_end > 0 and not steps[used_end - 1][0][0] < end:
        used_end -= 1
    past = order - (1 if implicit else 0)
    assert used_end >= past, "Insufficient past data."
    ids = [st[0] for st in steps[used_end - past:used_end]]
    if implicit:
        assert steps[used_end - 1][1] is not None, "Must have substep data for implicit stepping."
        ids.append(steps[used_end - 1][1])
    return ids


def _merge_to_small_steps(local, remote, local_implicit, remote_implicit, small_order,
                          small_implicit):
    """AdamsLts.cpp:214-277: the most recent values of the union of the control times, without
    the values that are only used for interpolation"""
    li, ri = len(local) - 1, len(remote) - 1
    if not small_implicit:
        # don't use implicit interpolation points for an explicit step
        if local_implicit:
            li -= 1
        if remote_implicit:
            ri -= 1
    elif local_implicit and remote_implicit:
        # two times after the small step, one from each side: one of them (if they differ)
        # belongs to a later small step
        if local[li] < remote[ri]:
            ri -= 1
        else:
            li -= 1
    out = []
    for _ in range(small_order):
        if li < 0:
            assert ri >= 0, "Ran out of data"
            t = remote[ri]
            ri -= 1
        elif ri < 0:
            t = local[li]
            li -= 1
        else:
            t = max(local[li], remote[ri])
            if local[li] == t:
                li -= 1
            if remote[ri] == t:
                ri -= 1
        out.append(t)
    return out[::-1]


def _as_entries(times):
    return [(Fr(t[0]), Fr(t[1])) if isinstance(t, tuple) else (Fr(t), Fr(0)) for t in times]


def lts_coefficients(local_times, remote_times, start, end, local_order, remote_order=None,
                     small_order=None, exact=False, local_implicit=False, remote_implicit=False,
                     small_implicit=False):
    """{(local id, remote id): coefficient} of the boundary contribution to the local side's
    step from `start` to `end` (AdamsLts.cpp:330-437).  Ids in insertion order: a number t is
    the step id at t, a pair (t, size) the substep (predictor) id of the step from t to
    t + size (needed by the implicit, Adams-Moulton, schemes); integers or Fractions.  In the
    result a step id is keyed by its time, a substep id by the pair."""
    remote_order = local_order if remote_order is None else remote_order
    small_order = local_order if small_order is None else small_order
    local_entries, remote_entries = _as_entries(local_times), _as_entries(remote_times)
    start, end = Fr(start), Fr(end)
    if start == end:
        return {}
    key = lambda en: en[0] if en[1] == 0 else en
    coefs = {}
    small_end = end
    while True:
        lids = _relevant(local_entries, small_end, local_order, local_implicit)
        rids = _relevant(remote_entries, small_end, remote_order, remote_implicit)
        if not coefs and (small_order, small_implicit) == (local_order, local_implicit) == \
                (remote_order, remote_implicit) and lids == rids:
            # no local time stepping at this boundary (lts_coefficients_for_gts)
            ct = [_exact_time(en) for en in lids]
            coefs = {(key(en), key(en)): c for en, c in zip(lids, _ab_exact(ct, start, end))}
            break
        lct = [_exact_time(en) for en in lids]
        rct = [_exact_time(en) for en in rids]
        small = _merge_to_small_steps(lct, rct, local_implicit, remote_implicit, small_order,
                                      small_implicit)
        current = small[-2] if small_implicit else small[-1]
        assert current >= start, "the start time is not a step boundary"
        small_coefs = _ab_exact(small, current, small_end)
        for m, tm in enumerate(small):
            li = _lagrange_exact(lct, tm)
            ri = _lagrange_exact(rct, tm)
            for a, la in zip(lids, li):
                if la == 0:
                    continue
                for b, rb in zip(rids, ri):
                    if rb == 0:
                        continue
                    k2 = (key(a), key(b))
                    coefs[k2] = coefs.get(k2, Fr(0)) + small_coefs[m] * la * rb
        if current == start:
            break
        small_end = current
    if exact:
        return dict(coefs)
    return {k: float(v) for k, v in coefs.items()}


# --------------------------------------------------------------------------------------
# evolution
# --------------------------------------------------------------------------------------
class LtsEvolution:
    """Fixed-level LTS evolution on one block-structured domain with conforming faces.

    levels [nelem]: element e takes steps dt_coarse / 2^levels[e].  Times are counted in
    ticks of the finest step.  past_states(j) -> u [nelem, C, n]: for every element its
    state at t0 - j * (its own step), j = 1 .. order - 1."""

    def __init__(self, system, N, invjac, static_fields, nbr, levels, order, t0, dt_coarse,
                 u0, past_states, gauge_params=orc.GAUGE_HARMONIC, ext_u=None, nbr_dir=None,
                 face_perm=None, static_face=None, mortars=None, post_update=None, coords=None):
        self.system, self.N, self.k = system, N, int(order)
        self.J, self.stat = invjac, static_fields
        self.nbr = np.asarray(nbr, dtype=np.int64)
        self.nelem = self.nbr.shape[0]
        self.levels = np.asarray(levels, dtype=np.int64)
        self.lmax = int(self.levels.max())
        self.stride = (2 ** (self.lmax - self.levels)).astype(np.int64)   # ticks per step
        self.tick_size = dt_coarse / 2 ** self.lmax
        self.t0 = t0
        self.gp, self.ext_u = gauge_params, ext_u
        self.coords = coords       # for Bjorhus external faces / coordinate-dependent gauges
        self.nbr_dir = nbr_dir if nbr_dir is not None else np.tile(
            np.array([1, 0, 3, 2, 5, 4]), (self.nelem, 1))
        self.face_perm = face_perm if face_perm is not None else np.zeros((self.nelem, 6), int)
        # the static fields that enter dg_package_data: SW gamma2; GH gamma1, gamma2
        self.static_face = static_fields if static_face is None else static_face
        # non-conforming (2:1) mortars between aligned blocks: rows (coarse element, direction,
        # fine element, direction, size_a, size_b) as in oracle.dg_rhs; both faces read HANGING
        self.mortars = [] if mortars is None else [tuple(int(v) for v in m) for m in mortars]
        assert all(m[3] >> 3 == 0 for m in self.mortars), "aligned mortars only"
        self.hanging = self.nbr == orc.HANGING
        if self.mortars:
            self.P = [np.eye(N)] + [orc.projection_matrix_parent_to_child(N, N, sz)
                                    for sz in (orc.MORTAR_LOWER_HALF, orc.MORTAR_UPPER_HALF)]
            self.R = [np.eye(N)] + [orc.projection_matrix_child_to_parent(N, N, sz)
                                    for sz in (orc.MORTAR_LOWER_HALF, orc.MORTAR_UPPER_HALF)]
        # external faces keep their boundary condition inside the "volume" part
        self.nbr_ext = np.where((self.nbr >= 0) | self.hanging, -1, self.nbr).astype(np.int32)
        # action after the step of an element (dg::Actions::Filter in the LTS action list)
        self.post = post_update if post_update is not None else (lambda v: v)
        self.u = u0.copy()
        self.tick = 0
        self.vol_hist = [[] for _ in range(self.nelem)]      # (tick, dt_u)
        self.face_hist = [[[] for _ in range(6)] for _ in range(self.nelem)]  # (tick, pk, mag)
        for j in range(self.k - 1, 0, -1):
            up = past_states(j)
            self._evaluate(np.arange(self.nelem), up, -j * self.stride)
        self.corrections_evaluated = 0

    def time(self, tick=None):
        return self.t0 + (self.tick if tick is None else tick) * self.tick_size

    # volume part (+ external boundary conditions) and the packaged data of the faces
    def _evaluate(self, elems, u_all, ticks):
        elems = np.asarray(elems)
        if len(elems) == 0:
            return
        ticks = np.broadcast_to(ticks, (self.nelem,)) if np.ndim(ticks) else np.full(
            self.nelem, ticks)
        dt = orc.dg_rhs(self.system, self.N, u_all[elems], self.J[elems], self.stat[elems],
                        self.nbr_ext[elems], gauge_params=self.gp, ext_u=self.ext_u,
                        coords=None if self.coords is None else self.coords[elems],
                        nbr_dir=np.ascontiguousarray(self.nbr_dir[elems], dtype=np.int32),
                        face_perm=np.ascontiguousarray(self.face_perm[elems], dtype=np.int32))
        for a, e in enumerate(elems):
            self.vol_hist[e].append((int(ticks[e]), dt[a]))
            for d in range(6):
                if self.nbr[e, d] >= 0 or self.hanging[e, d]:
                    pk, mag = orc.face_packaged_data(self.system, self.N, u_all[e], self.J[e],
                                                     self.static_face[e], d)
                    self.face_hist[e][d].append((int(ticks[e]), pk, mag))

    def _coupling(self, e, d, local, remote):
        """lifted dg_boundary_terms of the element's packaged data `local` and the
        neighbour's `remote` (compute_correction_coupling, ApplyBoundaryCorrections.hpp:
        814-905 with Gauss-Lobatto points)"""
        C = 5 if self.system == 0 else 50
        f = self.N * self.N
        omap = orc.orient_face_map(self.N, int(self.face_perm[e, d]))
        pk_ext = np.ascontiguousarray(remote[1][:, omap])
        pk_int = np.ascontiguousarray(local[1])
        corr = np.zeros((C, f))
        L = orc.lib()
        if self.system == 0:
            L.orc_sw_boundary_terms(f, orc._p(pk_int), orc._p(pk_ext), orc._p(corr))
        else:
            L.orc_gh_boundary_terms(f, orc._p(pk_int), orc._p(pk_ext), orc._p(corr))
        self.corrections_evaluated += 1
        return corr * (-0.5 * self.N * (self.N - 1) * local[2])     # LiftFlux.hpp:57-61

    def _to_mortar(self, pk, sa, sb):
        """project_to_mortar (MortarHelpers.hpp:74-101): first face dimension, then the
        second, of every packaged component (the characteristic speeds included)"""
        N = self.N
        x = pk.reshape(pk.shape[0], N, N)          # [c, b, a]
        x = np.einsum("ta,cba->cbt", self.P[sa], x)
        x = np.einsum("tb,cba->cta", self.P[sb], x)
        return np.ascontiguousarray(x.reshape(pk.shape[0], N * N))

    def _boundary_terms(self, pk_int, pk_ext):
        C = 5 if self.system == 0 else 50
        f = self.N * self.N
        corr = np.zeros((C, f))
        L = orc.lib()
        a, b = np.ascontiguousarray(pk_int), np.ascontiguousarray(pk_ext)
        if self.system == 0:
            L.orc_sw_boundary_terms(f, orc._p(a), orc._p(b), orc._p(corr))
        else:
            L.orc_gh_boundary_terms(f, orc._p(a), orc._p(b), orc._p(corr))
        self.corrections_evaluated += 1
        return corr

    def _mortar_coupling(self, m, fine_side, local, remote):
        """coupling across the mortar m for the fine element (its face is the mortar) or
        for the coarse one (correction on the mortar, project_from_mortar, lift on its face)"""
        _, _, _, _, sa, sb = m
        N = self.N
        if fine_side:
            corr = self._boundary_terms(local[1], self._to_mortar(remote[1], sa, sb))
            return corr * (-0.5 * N * (N - 1) * local[2])
        corr = self._boundary_terms(self._to_mortar(local[1], sa, sb), remote[1])
        x = corr.reshape(corr.shape[0], N, N)
        x = np.einsum("at,cbt->cba", self.R[sa], x)
        x = np.einsum("bt,cta->cba", self.R[sb], x)
        return x.reshape(corr.shape[0], N * N) * (-0.5 * N * (N - 1) * local[2])

    def _finalize(self, e, end_tick):
        """the step of element e that ends at end_tick: UpdateU with the volume history,
        then the boundary deltas of its internal faces"""
        k, s = self.k, int(self.stride[e])
        start = end_tick - s
        hist = self.vol_hist[e][-k:]
        assert hist[-1][0] == start and len(hist) == k
        coefs = orc.ab_coefficients_frac([Fr(h[0]) for h in hist], Fr(start), Fr(end_tick),
                                         self.tick_size)
        u = self.u[e].copy()
        for c, h in zip(coefs, hist):
            u += c * h[1]
        for d in range(6):
            nb = int(self.nbr[e, d])
            if nb < 0:
                continue
            nd = int(self.nbr_dir[e, d])
            local = self.face_hist[e][d][-k:]
            remote = [h for h in self.face_hist[nb][nd] if h[0] < end_tick]
            terms = lts_coefficients([h[0] for h in local], [h[0] for h in remote], start,
                                     end_tick, k)
            lifted = np.zeros_like(local[0][1][:u.shape[0]])
            by_tick_l = {h[0]: h for h in local}
            by_tick_r = {h[0]: h for h in remote}
            for (tl, tr), c in terms.items():
                lifted = lifted + (c * self.tick_size) * self._coupling(
                    e, d, by_tick_l[int(tl)], by_tick_r[int(tr)])
            u[:, orc.face_point_indices(self.N, d)] += lifted
        for m in self.mortars:
            ec, dc, ef, df = m[0], m[1], m[2], m[3] & 7
            for fine_side, (own, d, other, od) in ((False, (ec, dc, ef, df)),
                                                   (True, (ef, df, ec, dc))):
                if own != e:
                    continue
                local = self.face_hist[e][d][-k:]
                remote = [h for h in self.face_hist[other][od] if h[0] < end_tick]
                terms = lts_coefficients([h[0] for h in local], [h[0] for h in remote], start,
                                         end_tick, k)
                by_tick_l = {h[0]: h for h in local}
                by_tick_r = {h[0]: h for h in remote}
                lifted = 0.0
                for (tl, tr), c in terms.items():
                    lifted = lifted + (c * self.tick_size) * self._mortar_coupling(
                        m, fine_side, by_tick_l[int(tl)], by_tick_r[int(tr)])
                u[:, orc.face_point_indices(self.N, d)] += lifted
        return u

    def _prune(self):
        depth = self.k + int(self.stride.max() // self.stride.min()) + 1
        for e in range(self.nelem):
            del self.vol_hist[e][:-self.k]
            for d in range(6):
                del self.face_hist[e][d][:-depth]

    def take_ticks(self, n):
        """advance by n ticks of the finest step; the state is complete (all elements at
        the same time) whenever the tick count is a multiple of 2^(max level - min level)"""
        for _ in range(n):
            T = self.tick
            active = np.nonzero(T % self.stride == 0)[0]
            self._evaluate(active, self.u, T)
            ending = np.nonzero((T + 1) % self.stride == 0)[0]
            new = {e: self.post(self._finalize(e, T + 1)) for e in ending}
            for e, v in new.items():
                self.u[e] = v
            self.tick = T + 1
            self._prune()

    def take_coarse_steps(self, n):
        self.take_ticks(n * int(self.stride.max()))
