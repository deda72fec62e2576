// Local time stepping with Adams-Bashforth: host logic.
//   dgrhs_adams_lts_coefficients  TimeSteppers::adams_lts::lts_coefficients for explicit
//                                 schemes (src/Time/TimeSteppers/AdamsLts.cpp:165-437)
//   dgrhs_lts_*                   elements with step sizes dt_coarse / 2^level: the volume
//                                 part of the time derivative (with the external boundary
//                                 conditions, as ComputeTimeDerivative applies them) goes
//                                 through the element's own Adams-Bashforth history
//                                 (Actions/UpdateU.hpp:44-120); the boundary corrections of
//                                 internal faces are integrated over the step from the
//                                 histories of both sides (AdamsBashforth.cpp:264-281,
//                                 ApplyBoundaryCorrections.hpp:797-1010 with
//                                 local_time_stepping == true)
// Schedule (one stream; the reference's asynchronous elements obey the same dependencies):
// time runs in ticks of the finest step.  At tick T the elements with a step boundary at T
// -- a suffix of the element order, elements are sorted by level -- are evaluated (volume
// part into the history slot of their level, faces into their snapshot ring); then the
// elements whose step ends at T + 1 are completed: u += sum_i c_i dt_i (volume history),
// u += sum_ij c_ij lift(D(local_i, remote_j)) per internal face.  All remote values a
// completing element needs have ticks <= T and are in the rings by then.
#include <cmath>
#include <cstring>
#include <map>
#include <tuple>

#include "ctx.cuh"

namespace {

struct Term {
  int li, ri;  // indices into the local / remote id lists
  double coef;
};

using Ticks = std::vector<long long>;

// an id of a boundary history: the step id at `tick` (sub == 0) or the substep (predictor) id
// of the step from tick to tick + sub; adams_lts::exact_substep_time (AdamsLts.cpp:29-38)
struct Id {
  long long tick, sub;
  long long time() const { return tick + sub; }
  bool operator==(const Id& o) const { return tick == o.tick && sub == o.sub; }
};
struct Scheme {
  bool implicit;
  int order;
  bool operator==(const Scheme& o) const { return implicit == o.implicit && order == o.order; }
};

// AdamsLts.cpp:173-206 (find_relevant_ids): by position; indices into `ids`
bool relevant(const std::vector<Id>& ids, long long end, const Scheme& sch,
              std::vector<int>* idx, const char** why) {
  std::vector<int> steps, substep_of;  // step entries and the substep entry of each (or -1)
  for (int i = 0; i < (int)ids.size(); ++i) {
    if (ids[i].sub == 0) {
      steps.push_back(i);
      substep_of.push_back(-1);
    } else {
      if (steps.empty() || ids[steps.back()].tick != ids[i].tick) {
        *why = "substep id without its step id";
        return false;
      }
      substep_of.back() = i;
    }
  }
  int used_end = (int)steps.size();
  while (used_end > 0 && !(ids[steps[used_end - 1]].tick < end)) --used_end;
  const int past = sch.order - (sch.implicit ? 1 : 0);
  if (used_end < past) {
    *why = "Insufficient past data.";
    return false;
  }
  idx->clear();
  for (int i = used_end - past; i < used_end; ++i) idx->push_back(steps[i]);
  if (sch.implicit) {
    if (used_end < 1 || substep_of[used_end - 1] < 0) {
      *why = "Must have substep data for implicit stepping.";
      return false;
    }
    idx->push_back(substep_of[used_end - 1]);
  }
  return true;
}

// AdamsLts.cpp:280-305
std::vector<double> interpolation_coefficients(const Ticks& control, long long t, double origin,
                                               double tick_size) {
  std::vector<double> c(control.size(), 0.0);
  for (size_t i = 0; i < control.size(); ++i)
    if (control[i] == t) {
      c[i] = 1.0;
      return c;
    }
  std::vector<double> fp(control.size());
  for (size_t i = 0; i < control.size(); ++i) fp[i] = origin + (double)control[i] * tick_size;
  const double x = origin + (double)t * tick_size;
  for (size_t j = 0; j < control.size(); ++j) {
    double r = 1.0;
    for (size_t m = 0; m < control.size(); ++m)
      if (m != j) r *= (fp[m] - x) / (fp[m] - fp[j]);
    c[j] = r;
  }
  return c;
}

// lts_coefficients (AdamsLts.cpp:330-437); false: insufficient data
bool lts_coefficients(const std::vector<Id>& local, const std::vector<Id>& remote,
                      long long start, long long end, const Scheme& lo, const Scheme& ro,
                      const Scheme& so, double origin, double tick_size, std::vector<Term>* out,
                      const char** why) {
  out->clear();
  if (start == end) return true;
  std::vector<std::tuple<int, int, double>> raw;
  long long small_end = end;
  for (;;) {
    std::vector<int> li, ri;
    if (!relevant(local, small_end, lo, &li, why) || !relevant(remote, small_end, ro, &ri, why))
      return false;
    Ticks lt, rt;
    bool same_ids = li.size() == ri.size();
    for (int i : li) lt.push_back(local[i].time());
    for (int i : ri) rt.push_back(remote[i].time());
    for (size_t i = 0; same_ids && i < li.size(); ++i) same_ids = local[li[i]] == remote[ri[i]];
    if (raw.empty() && so == lo && so == ro && same_ids) {
      // the sides step at the same rate: lts_coefficients_for_gts
      const auto g = dgrhs_internal_ab_coefficients_ticks(lt, start, end, tick_size);
      for (size_t s = 0; s < g.size(); ++s) out->push_back({li[s], ri[s], g[s]});
      return true;
    }
    // merge_to_small_steps (AdamsLts.cpp:214-277)
    Ticks small((size_t)so.order);
    {
      int a = (int)lt.size() - 1, b = (int)rt.size() - 1;
      if (!so.implicit) {
        // don't use implicit interpolation points for an explicit step
        if (lo.implicit) --a;
        if (ro.implicit) --b;
      } else if (lo.implicit && ro.implicit) {
        // of the two times after the small step one belongs to a later small step
        if (lt[a] < rt[b])
          --b;
        else
          --a;
      }
      for (int o = so.order - 1; o >= 0; --o) {
        if (a < 0) {
          if (b < 0) {
            *why = "Ran out of data";
            return false;
          }
          small[o] = rt[b--];
        } else if (b < 0) {
          small[o] = lt[a--];
        } else {
          small[o] = std::max(lt[a], rt[b]);
          const bool la = lt[a] == small[o], rb = rt[b] == small[o];
          if (la) --a;
          if (rb) --b;
        }
      }
    }
    if (so.implicit && small.size() < 2) {
      *why = "implicit small-step scheme of order 1";
      return false;
    }
    const long long current = small[small.size() - (so.implicit ? 2 : 1)];
    if (current < start) {
      *why = "the start time is not a step boundary";
      return false;
    }
    const auto sc = dgrhs_internal_ab_coefficients_ticks(small, current, small_end, tick_size);
    for (size_t m = 0; m < small.size(); ++m) {
      const auto lc = interpolation_coefficients(lt, small[m], origin, tick_size);
      const auto rc = interpolation_coefficients(rt, small[m], origin, tick_size);
      for (size_t a = 0; a < lc.size(); ++a) {
        if (lc[a] == 0.0) continue;
        for (size_t b = 0; b < rc.size(); ++b) {
          if (rc[b] == 0.0) continue;
          raw.emplace_back(li[a], ri[b], sc[m] * lc[a] * rc[b]);
        }
      }
    }
    if (current == start) break;
    small_end = current;
  }
  // combine duplicate entries, sorted like the reference's (local id, remote id): a step id
  // before its substep id
  const auto key = [&](const std::tuple<int, int, double>& x) {
    const Id &a = local[std::get<0>(x)], &b = remote[std::get<1>(x)];
    return std::make_tuple(a.tick, a.sub != 0, b.tick, b.sub != 0);
  };
  std::stable_sort(raw.begin(), raw.end(),
                   [&](const auto& x, const auto& y) { return key(x) < key(y); });
  for (const auto& t : raw) {
    if (!out->empty() && out->back().li == std::get<0>(t) && out->back().ri == std::get<1>(t))
      out->back().coef += std::get<2>(t);
    else
      out->push_back({std::get<0>(t), std::get<1>(t), std::get<2>(t)});
  }
  return true;
}

std::vector<Id> step_ids(const Ticks& ticks) {
  std::vector<Id> ids;
  for (long long t : ticks) ids.push_back({t, 0});
  return ids;
}

struct LtsState {
  int order = 0, nlevels = 0, depth = 0, max_terms = 0;
  std::vector<int> level_begin;        // [nlevels + 1] element ranges, coarse first
  std::vector<long long> stride;       // ticks per step of each level
  double t0 = 0.0, tick_size = 0.0;
  long long tick = 0;
  int past_set = 0;                    // bit j: past state j given
  int32_t* level_dev = nullptr;        // [E]
  int32_t* nbr_ext = nullptr;          // neighbour table with the internal faces masked
  double* fh = nullptr;                // [E][depth][6][C][f]
  double* acc = nullptr;               // [E][6][C][f]
  dg::LtsTerm* terms_dev = nullptr;    // [nlevels][max_terms]
  std::vector<double*> vol;            // [order] full-state derivative buffers
  std::vector<std::vector<bool>> adjacent;  // level pairs that share a face
  // mode 1: faces between elements of the same level are evaluated like GTS faces (their
  // corrections enter the volume history) and the stepper update is fused into the volume
  // kernel: the state of a level then alternates between the two state buffers
  int same_level_in_volume = 1;
  bool fused = false;
  double* buf[2] = {nullptr, nullptr};
  std::vector<int> parity;             // [nlevels] which buffer holds the level's state
  uint8_t* in_history = nullptr;         // [E][6] faces kept in the boundary histories
  uint8_t* mortar_in_history = nullptr;  // [n_mortars] (order of the context's mortar table)
  int n_groups = 0;
  std::vector<char> level_has_conforming, level_is_coarse_side, level_is_fine_side;
};

inline int mod(long long a, int m) { return (int)(((a % m) + m) % m); }

LtsState* state(dgrhs_ctx* c) { return static_cast<LtsState*>(c->lts); }

// Adams-Bashforth coefficients of the step m -> m + 1 of `level` (uniform history)
std::vector<double> step_coefficients(const LtsState* s, int level, long long m) {
  Ticks ticks;
  const long long st = s->stride[level];
  for (int i = s->order - 1; i >= 0; --i) ticks.push_back((m - i) * st);
  return dgrhs_internal_ab_coefficients_ticks(ticks, m * st, (m + 1) * st, s->tick_size);
}

// volume part + face snapshots of the elements of `level` at step index m; with `update` (and
// the fused mode) the volume kernel also writes u + sum_i c_i dt_i into the level's other
// state buffer (UpdateU; the boundary deltas follow when the step completes)
int evaluate_level(dgrhs_ctx* c, LtsState* s, int level, long long m, bool update) {
  const DgNOps* ops = dgrhs_nops(c->N);
  const int eb = s->level_begin[level], ee = s->level_begin[level + 1];
  if (ee <= eb) return 0;
  const int k = s->order;
  double* const keep = c->u;
  if (s->fused) c->u = s->buf[s->parity[level]];
  int rc = ops->lts_snapshot(c, s->fh, s->in_history, s->depth, mod(m, s->depth), eb, ee);
  if (!rc && update && s->fused) {
    const auto coef = step_coefficients(s, level, m);
    dg::UpdateArgs up{};
    up.u_new = s->buf[1 - s->parity[level]];
    up.a = 1.0;
    up.nterms = k - 1;
    for (int j = 0; j < k - 1; ++j) {
      up.c[j] = coef[j];
      up.v[j] = s->vol[mod(m - (k - 1) + j, k)];
    }
    up.c_new = coef[k - 1];
    rc = ops->lts_evaluate(c, s->nbr_ext, s->mortar_in_history, s->vol[mod(m, k)], eb, ee, &up);
    s->parity[level] ^= 1;
  } else if (!rc) {
    rc = ops->lts_evaluate(c, s->nbr_ext, s->mortar_in_history, s->vol[mod(m, k)], eb, ee,
                           nullptr);
  }
  c->u = keep;
  return rc;
}

// the step of `level` from step index m to m + 1
int complete_level(dgrhs_ctx* c, LtsState* s, int level, long long m) {
  const DgNOps* ops = dgrhs_nops(c->N);
  const int eb = s->level_begin[level], ee = s->level_begin[level + 1];
  if (ee <= eb) return 0;
  const int k = s->order;
  const long long st = s->stride[level], start = m * st, end = start + st;
  double* const u_level = s->fused ? s->buf[s->parity[level]] : c->u;
  // UpdateU with the element's own history (oldest term first, like the GTS update); in the
  // fused mode the volume kernel did it when the step was evaluated
  if (!s->fused) {
    std::vector<const double*> v;
    const size_t off = (size_t)eb * c->C * c->npad;
    for (int i = k - 1; i >= 0; --i) v.push_back(s->vol[mod(m - i, k)] + off);
    const auto coef = step_coefficients(s, level, m);
    if (dgrhs_internal_lincomb_range(c, u_level + off, 1.0, coef, v,
                                     (size_t)(ee - eb) * c->C * c->npad))
      return 1;
  }
  bool any = false;
  for (int nl = 0; nl < s->nlevels; ++nl)
    if (s->adjacent[level][nl] && !(s->same_level_in_volume && nl == level)) any = true;
  if (!any) return 0;
  // boundary deltas: one coefficient list per level of the neighbour
  dg::LtsBoundaryArgs a{};
  a.fh = s->fh;
  a.invjac = c->invjac;
  a.stat = c->stat;
  a.nbr = c->nbr;
  a.nbr_face = c->nbr_face;
  a.level = s->level_dev;
  a.terms = s->terms_dev;
  a.max_terms = s->max_terms;
  a.depth = s->depth;
  a.elem_begin = eb;
  a.elem_end = ee;
  a.acc = s->acc;
  a.u = u_level;
  a.in_history = s->in_history;
  std::vector<dg::LtsTerm> host((size_t)s->nlevels * s->max_terms);
  Ticks local;
  for (int i = k - 1; i >= 0; --i) local.push_back((m - i) * st);
  for (int nl = 0; nl < s->nlevels; ++nl) {
    a.nterms[nl] = 0;
    if (!s->adjacent[level][nl] || (s->same_level_in_volume && nl == level)) continue;
    // the neighbour's snapshots before `end`: its step indices up to the last one that
    // starts before `end`, as far back as the ring holds them
    const long long sn = s->stride[nl];
    const long long last = (end - 1 >= 0) ? (end - 1) / sn : -((-(end - 1) + sn - 1) / sn);
    Ticks remote;
    std::vector<long long> remote_m;
    for (long long j = last - (s->depth - 1); j <= last; ++j) {
      remote.push_back(j * sn);
      remote_m.push_back(j);
    }
    std::vector<Term> terms;
    const char* why = "";
    const Scheme ab{false, k};
    if (!lts_coefficients(step_ids(local), step_ids(remote), start, end, ab, ab, ab, s->t0,
                          s->tick_size, &terms, &why))
      return fail("LTS coefficients (levels %d / %d): %s", level, nl, why);
    if ((int)terms.size() > s->max_terms)
      return fail("internal error: %d LTS terms, room for %d", (int)terms.size(), s->max_terms);
    for (size_t t = 0; t < terms.size(); ++t) {
      // a snapshot older than the ring would alias a newer one
      if (remote_m[terms[t].ri] <= last - s->depth)
        return fail("internal error: LTS snapshot ring too short");
      host[(size_t)nl * s->max_terms + t] = {mod(m - (k - 1) + terms[t].li, s->depth),
                                             mod(remote_m[terms[t].ri], s->depth),
                                             terms[t].coef};
    }
    a.nterms[nl] = (int)terms.size();
  }
  CU(cudaMemcpyAsync(s->terms_dev, host.data(), host.size() * sizeof(dg::LtsTerm),
                     cudaMemcpyHostToDevice, c->stream));
  // (pageable source: the copy is staged before the call returns)
  if (s->level_has_conforming[level] && ops->lts_boundary(c, &a)) return 1;
  if (s->level_is_coarse_side[level] || s->level_is_fine_side[level]) {
    dg::LtsMortarArgs ma{};
    ma.fh = s->fh;
    ma.invjac = c->invjac;
    ma.stat = c->stat;
    ma.faces = c->mortar_faces;
    ma.mortars = c->mortar_table;
    ma.mortar_in_history = s->mortar_in_history;
    ma.P = c->mortar_P;
    ma.R = c->mortar_R;
    ma.level = s->level_dev;
    ma.terms = s->terms_dev;
    for (int nl = 0; nl < dg::kLtsMaxLevels; ++nl) ma.nterms[nl] = a.nterms[nl];
    ma.max_terms = s->max_terms;
    ma.depth = s->depth;
    ma.elem_begin = eb;
    ma.elem_end = ee;
    ma.acc = s->acc;
    for (int role = 0; role < 2; ++role) {
      if (!(role == 0 ? s->level_is_coarse_side[level] : s->level_is_fine_side[level])) continue;
      ma.role = role;
      if (ops->lts_mortar(c, &ma, s->n_groups)) return 1;
    }
  }
  a.terms = nullptr;   // u += acc on the faces in the histories
  return ops->lts_boundary(c, &a);
}

// the step of `level` m -> m + 1 with the action that follows it in the reference's LTS
// action list: dg::Actions::Filter on the elements that took the step
int complete_level_and_filter(dgrhs_ctx* c, LtsState* s, int level, long long m) {
  if (complete_level(c, s, level, m)) return 1;
  if (!c->filterF) return 0;
  const int eb = s->level_begin[level], ee = s->level_begin[level + 1];
  double* const u_level = s->fused ? s->buf[s->parity[level]] : c->u;
  return dgrhs_nops(c->N)->filter_range(c, u_level + (size_t)eb * c->C * c->npad,
                                        (ee - eb) * c->C);
}

}  // namespace

void dgrhs_internal_lts_free(dgrhs_ctx* c) {
  LtsState* s = state(c);
  if (!s) return;
  for (double* p : s->vol) cudaFree(p);
  if (s->level_dev) cudaFree(s->level_dev);
  if (s->nbr_ext) cudaFree(s->nbr_ext);
  if (s->fh) cudaFree(s->fh);
  if (s->acc) cudaFree(s->acc);
  if (s->terms_dev) cudaFree(s->terms_dev);
  if (s->in_history) cudaFree(s->in_history);
  if (s->mortar_in_history) cudaFree(s->mortar_in_history);
  delete s;
  c->lts = nullptr;
}

extern "C" {

int dgrhs_adams_lts_coefficients_general(
    int local_implicit, int local_order, int remote_implicit, int remote_order,
    int small_step_implicit, int small_step_order, int n_local, const long long* local_ticks,
    const long long* local_substep_sizes, int n_remote, const long long* remote_ticks,
    const long long* remote_substep_sizes, long long start_tick, long long end_tick,
    double time_origin, double tick_size, int max_terms, int* n_terms, int* local_index,
    int* remote_index, double* coefficients) {
  for (int o : {local_order, remote_order, small_step_order})
    if (o < 1 || o > 8) return fail("order must be in [1, 8]");
  if (n_local < 1 || n_remote < 1) return fail("empty history");
  std::vector<Id> local, remote;
  for (int i = 0; i < n_local; ++i)
    local.push_back({local_ticks[i], local_substep_sizes ? local_substep_sizes[i] : 0});
  for (int i = 0; i < n_remote; ++i)
    remote.push_back({remote_ticks[i], remote_substep_sizes ? remote_substep_sizes[i] : 0});
  std::vector<Term> terms;
  const char* why = "";
  if (!lts_coefficients(local, remote, start_tick, end_tick,
                        Scheme{local_implicit != 0, local_order},
                        Scheme{remote_implicit != 0, remote_order},
                        Scheme{small_step_implicit != 0, small_step_order}, time_origin,
                        tick_size, &terms, &why))
    return fail("%s", why);
  if ((int)terms.size() > max_terms)
    return fail("%d terms, room for %d", (int)terms.size(), max_terms);
  *n_terms = (int)terms.size();
  for (size_t t = 0; t < terms.size(); ++t) {
    local_index[t] = terms[t].li;
    remote_index[t] = terms[t].ri;
    coefficients[t] = terms[t].coef;
  }
  return 0;
}

int dgrhs_adams_lts_coefficients(int local_order, int remote_order, int small_step_order,
                                 int n_local, const long long* local_ticks, int n_remote,
                                 const long long* remote_ticks, long long start_tick,
                                 long long end_tick, double time_origin, double tick_size,
                                 int max_terms, int* n_terms, int* local_index,
                                 int* remote_index, double* coefficients) {
  return dgrhs_adams_lts_coefficients_general(
      0, local_order, 0, remote_order, 0, small_step_order, n_local, local_ticks, nullptr,
      n_remote, remote_ticks, nullptr, start_tick, end_tick, time_origin, tick_size, max_terms,
      n_terms, local_index, remote_index, coefficients);
}

int dgrhs_lts_init(dgrhs_ctx* c, int order, double t0, double dt_coarse, const int32_t* levels) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (order < 1 || order > 8) return fail("order must be in [1, 8]");
  if (!(dt_coarse > 0.0)) return fail("time step must be positive");
  if (c->n_send > 0 || c->nccl_comm)
    return fail("local time stepping runs on one GPU (no halo exchange)");
  if (c->n_pmortar_faces > 0) return fail("local time stepping: p-mortars are not supported");
  if (c->n_mortar_faces != c->n_mortar_faces_local)
    return fail("local time stepping: mortars with a remote side are not supported");
  if (c->mesh_v) return fail("local time stepping on a moving mesh is not supported");
  if (c->gauge == DGRHS_GAUGE_ANALYTIC_GAUGE_WAVE)
    return fail("local time stepping: time-dependent gauge fields are not supported");
  if (c->nbr_host.empty()) return fail("dgrhs_lts_init needs dgrhs_set_geometry first");
  if (!c->nbr_face && !c->aligned_table_ok)
    return fail("neighbor table is not that of aligned blocks: call "
                "dgrhs_set_neighbor_orientations");
  dgrhs_internal_lts_free(c);
  auto* s = new LtsState;
  c->lts = s;
  s->order = order;
  int lmax = 0;
  for (int e = 0; e < c->nelem; ++e) {
    if (levels[e] < 0 || levels[e] >= dg::kLtsMaxLevels)
      return fail("step-size level of element %d out of range [0, %d)", e, dg::kLtsMaxLevels);
    if (e > 0 && levels[e] < levels[e - 1])
      return fail("elements must be sorted by step-size level (coarse steps first)");
    lmax = std::max(lmax, levels[e]);
  }
  s->nlevels = lmax + 1;
  s->level_begin.assign(s->nlevels + 1, c->nelem);
  for (int l = 0; l <= s->nlevels; ++l)
    s->level_begin[l] = (int)(std::lower_bound(levels, levels + c->nelem, l) - levels);
  s->stride.resize(s->nlevels);
  for (int l = 0; l < s->nlevels; ++l) s->stride[l] = 1LL << (lmax - l);
  s->t0 = t0;
  s->tick_size = dt_coarse / (double)(1LL << lmax);
  s->tick = 0;
  // which levels meet at a face, the largest step ratio across a face, and the faces /
  // mortars that are kept in the boundary histories
  s->same_level_in_volume = c->lts_mode != 0;
  s->adjacent.assign(s->nlevels, std::vector<bool>(s->nlevels, false));
  s->level_has_conforming.assign(s->nlevels, 0);
  s->level_is_coarse_side.assign(s->nlevels, 0);
  s->level_is_fine_side.assign(s->nlevels, 0);
  long long ratio = 1;
  std::vector<uint8_t> hist((size_t)c->nelem * 6, 0);
  const auto in_hist = [&](int e, int nb) {
    return !(s->same_level_in_volume && levels[e] == levels[nb]);
  };
  for (int e = 0; e < c->nelem; ++e)
    for (int d = 0; d < 6; ++d) {
      const int nb = c->nbr_host[(size_t)e * 6 + d];
      if (nb < 0) continue;
      if (nb >= c->nelem) return fail("local time stepping: ghost elements are not supported");
      s->adjacent[levels[e]][levels[nb]] = true;
      ratio = std::max(ratio, 1LL << std::abs(levels[e] - levels[nb]));
      if (in_hist(e, nb)) {
        hist[(size_t)e * 6 + d] = 1;
        s->level_has_conforming[levels[e]] = 1;
      }
    }
  s->n_groups = (int)c->mortar_faces_host.size() / 4;
  const int n_mortars = (int)c->mortar_table_host.size() / 4;
  std::vector<uint8_t> mhist(std::max(n_mortars, 1), 0);
  for (int g = 0; g < s->n_groups; ++g) {
    const int32_t* fc = c->mortar_faces_host.data() + 4 * (size_t)g;
    const int ec = fc[0], dc = fc[1];
    for (int m = fc[2]; m < fc[2] + fc[3]; ++m) {
      const int32_t* mt = c->mortar_table_host.data() + 4 * (size_t)m;
      const int ef = mt[0], df = mt[1] & 7;
      if (ec < 0 || ef < 0) return fail("local time stepping: mortars with a remote side");
      s->adjacent[levels[ec]][levels[ef]] = true;
      s->adjacent[levels[ef]][levels[ec]] = true;
      ratio = std::max(ratio, 1LL << std::abs(levels[ec] - levels[ef]));
      if (in_hist(ec, ef)) {
        mhist[m] = 1;
        hist[(size_t)ec * 6 + dc] = hist[(size_t)ef * 6 + df] = 1;
        s->level_is_coarse_side[levels[ec]] = 1;
        s->level_is_fine_side[levels[ef]] = 1;
      }
    }
  }
  // a coarse element completes its step with the fine neighbour's last order - 1 + ratio
  // snapshots; one more slot so that the neighbour's next evaluation does not overwrite
  s->depth = order + (int)ratio;
  s->max_terms = order * (order + (int)ratio);
  const size_t f = (size_t)c->N * c->N;
  if (dev_alloc(&s->fh, (size_t)c->nelem * s->depth * 6 * c->C * f)) return 1;
  if (dev_alloc(&s->acc, (size_t)c->nelem * 6 * c->C * f)) return 1;
  if (dev_alloc(&s->level_dev, (size_t)c->nelem)) return 1;
  if (dev_alloc(&s->nbr_ext, (size_t)c->nelem * 6)) return 1;
  if (dev_alloc(&s->terms_dev, (size_t)s->nlevels * s->max_terms)) return 1;
  for (int i = 0; i < order; ++i) {
    double* p = nullptr;
    if (dev_alloc(&p, c->state_len())) return 1;
    s->vol.push_back(p);
  }
  s->fused = s->same_level_in_volume && c->fuse_update && order <= 4;
  s->parity.assign(s->nlevels, 0);
  if (s->fused) {
    if (!c->u_alt && dev_alloc(&c->u_alt, c->state_len())) return 1;
    s->buf[0] = c->u;
    s->buf[1] = c->u_alt;
  }
  // the faces the face kernel does not evaluate read "no correction"
  std::vector<int32_t> ext(c->nbr_host);
  for (int e = 0; e < c->nelem; ++e)
    for (int d = 0; d < 6; ++d) {
      int32_t& v = ext[(size_t)e * 6 + d];
      if (v >= 0 && !(s->same_level_in_volume && levels[v] == levels[e]))
        v = dg::kLtsHistoryFace;
    }
  CU(h2d_table(s->nbr_ext, ext.data(), ext.size() * 4));
  CU(h2d_table(s->level_dev, levels, (size_t)c->nelem * 4));
  if (dev_alloc(&s->in_history, hist.size())) return 1;
  if (dev_alloc(&s->mortar_in_history, mhist.size())) return 1;
  CU(h2d_table(s->in_history, hist.data(), hist.size()));
  CU(h2d_table(s->mortar_in_history, mhist.data(), mhist.size()));
  // correction slots of faces in the histories are never written by the face / mortar kernels
  // of an evaluation: they must read zero (the volume kernel adds all six slots)
  CU(cudaMemsetAsync(c->corr, 0, (size_t)c->nelem * 6 * c->C * c->N * c->N * sizeof(double),
                     c->stream));
  // the tables above went through the legacy default stream, which the context's
  // non-blocking stream does not wait for: everything is in place before the first kernel
  CU(cudaDeviceSynchronize());
  return 0;
}

int dgrhs_lts_set_past_state(dgrhs_ctx* c, int j, const double* u_past) {
  CHECK_CTX(c);
  LtsState* s = state(c);
  if (!s) return fail("dgrhs_lts_init first");
  if (j < 1 || j >= s->order) return fail("past state index must be in [1, order)");
  if (s->tick != 0) return fail("past states belong to the start of the evolution");
  CU(cudaSetDevice(c->device));
  // evaluate on a scratch copy of the state: c->u holds the initial data
  double* keep = nullptr;
  if (dev_alloc(&keep, c->state_len())) return 1;
  CU(cudaMemcpyAsync(keep, c->u, c->state_len() * 8, cudaMemcpyDeviceToDevice, c->stream));
  int rc = dgrhs_internal_upload(c, c->u, u_past, c->C);
  const DgNOps* ops = dgrhs_nops(c->N);
  if (!rc) rc = ops->gauge(c, s->t0);
  for (int l = 0; l < s->nlevels && !rc; ++l) rc = evaluate_level(c, s, l, -(long long)j, false);
  if (cudaMemcpyAsync(c->u, keep, c->state_len() * 8, cudaMemcpyDeviceToDevice, c->stream) !=
      cudaSuccess)
    rc = 1;
  cudaStreamSynchronize(c->stream);
  cudaFree(keep);
  if (!rc) s->past_set |= 1 << j;
  return rc;
}

int dgrhs_lts_take_ticks(dgrhs_ctx* c, long long n_ticks) {
  CHECK_CTX(c);
  LtsState* s = state(c);
  if (!s) return fail("dgrhs_lts_init first");
  CU(cudaSetDevice(c->device));
  for (int j = 1; j < s->order; ++j)
    if (!(s->past_set & (1 << j)))
      return fail("past state %d of the Adams-Bashforth history is missing "
                  "(dgrhs_lts_set_past_state)", j);
  const DgNOps* ops = dgrhs_nops(c->N);
  for (long long it = 0; it < n_ticks; ++it) {
    const long long T = s->tick;
    if (ops->gauge(c, s->t0 + (double)T * s->tick_size)) return 1;
    ++c->rhs_evals;
    for (int l = 0; l < s->nlevels; ++l)
      if (T % s->stride[l] == 0 && evaluate_level(c, s, l, T / s->stride[l], true)) return 1;
    for (int l = 0; l < s->nlevels; ++l)
      if ((T + 1) % s->stride[l] == 0 &&
          complete_level_and_filter(c, s, l, (T + 1) / s->stride[l] - 1))
        return 1;
    s->tick = T + 1;
  }
  // the caller reads the state from c->u: bring the levels that sit in the other buffer back
  if (s->fused)
    for (int l = 0; l < s->nlevels; ++l) {
      if (s->parity[l] == 0) continue;
      const size_t off = (size_t)s->level_begin[l] * c->C * c->npad;
      const size_t len = (size_t)(s->level_begin[l + 1] - s->level_begin[l]) * c->C * c->npad;
      CU(cudaMemcpyAsync(s->buf[0] + off, s->buf[1] + off, len * 8, cudaMemcpyDeviceToDevice,
                         c->stream));
      s->parity[l] = 0;
    }
  return 0;
}

int dgrhs_lts_set_mode(dgrhs_ctx* c, int same_level_faces_in_volume_history) {
  CHECK_CTX(c);
  c->lts_mode = same_level_faces_in_volume_history ? 1 : 0;
  return 0;
}

int dgrhs_lts_ticks_per_coarse_step(dgrhs_ctx* c, long long* n) {
  CHECK_CTX(c);
  LtsState* s = state(c);
  if (!s) return fail("dgrhs_lts_init first");
  *n = s->stride[0];
  return 0;
}

int dgrhs_lts_time(dgrhs_ctx* c, double* time, long long* tick) {
  CHECK_CTX(c);
  LtsState* s = state(c);
  if (!s) return fail("dgrhs_lts_init first");
  if (time) *time = s->t0 + (double)s->tick * s->tick_size;
  if (tick) *tick = s->tick;
  return 0;
}

}  // extern "C"
