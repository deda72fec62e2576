// Host side of the C-ABI in include/dgrhs.h: context, kernel launchers,
// spectral matrices, time steppers (Adams-Bashforth with the reference's
// forward self-start, Rk3HesthavenSsp) and the single-operator entry points.
// Reference citations are in include/dgrhs.h next to each entry point.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "ctx.cuh"

namespace {

thread_local std::string g_error;
int64_t g_launches = 0;

// ---------------------------------------------------------------------------
// Spectral quantities: Legendre-Gauss-Lobatto nodes/weights (Kopriva Alg. 25,
// reference Legendre.cpp:187-232), barycentric weights (Alg. 30, Spectral.cpp:
// 84-104), differentiation matrix (Spectral.cpp:431-445).
// ---------------------------------------------------------------------------
void q_and_L(int deg, double x, double* q, double* L) {
  double Lm2 = 1.0, Lm1 = x, Ln = x;
  for (int k = 2; k <= deg; ++k) {
    Ln = ((2.0 * k - 1.0) * x * Lm1 - (k - 1.0) * Lm2) / k;
    Lm2 = Lm1;
    Lm1 = Ln;
  }
  const int k = deg + 1;
  const double Lp1 = ((2.0 * k - 1.0) * x * Ln - (k - 1.0) * Lm2) / k;
  *q = Lp1 - Lm2;
  *L = Ln;
}

void lgl(int num_points, std::vector<double>& x, std::vector<double>& w) {
  const int deg = num_points - 1;
  x.assign(num_points, 0.0);
  w.assign(num_points, 0.0);
  if (deg == 1) {
    x[0] = -1.0;
    x[1] = 1.0;
    w[0] = w[1] = 1.0;
    return;
  }
  x[0] = -1.0;
  x[deg] = 1.0;
  w[0] = w[deg] = 2.0 / (deg * (deg + 1.0));
  for (int j = 1; j < (deg + 1) / 2; ++j) {
    double lo = -cos((j - 0.25) * M_PI / deg - 0.375 / (deg * M_PI * (j - 0.25)));
    double hi = -cos((j + 0.75) * M_PI / deg - 0.375 / (deg * M_PI * (j + 0.75)));
    double flo, fhi, L;
    q_and_L(deg, lo, &flo, &L);
    q_and_L(deg, hi, &fhi, &L);
    for (int it = 0; it < 200; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (mid == lo || mid == hi) break;
      double fm;
      q_and_L(deg, mid, &fm, &L);
      if (fm == 0.0) {
        lo = hi = mid;
        break;
      }
      if ((fm < 0) == (flo < 0)) {
        lo = mid;
        flo = fm;
      } else {
        hi = mid;
        fhi = fm;
      }
    }
    const double root = 0.5 * (lo + hi);
    double q;
    q_and_L(deg, root, &q, &L);
    x[j] = root;
    x[deg - j] = -root;
    w[j] = w[deg - j] = 2.0 / (deg * (deg + 1.0) * L * L);
  }
  if (deg % 2 == 0) {
    double q, L;
    q_and_L(deg, 0.0, &q, &L);
    x[deg / 2] = 0.0;
    w[deg / 2] = 2.0 / (deg * (deg + 1.0) * L * L);
  }
}

void diff_matrix(int N, std::vector<double>& D) {
  std::vector<double> x, w;
  lgl(N, x, w);
  std::vector<double> bw(N, 1.0);
  for (int j = 1; j < N; ++j)
    for (int k = 0; k < j; ++k) {
      bw[k] *= x[k] - x[j];
      bw[j] *= x[j] - x[k];
    }
  for (int j = 0; j < N; ++j) bw[j] = 1.0 / bw[j];
  D.assign((size_t)N * N, 0.0);
  for (int i = 0; i < N; ++i) {
    double diag = 0.0;
    for (int j = 0; j < N; ++j)
      if (i != j) {
        D[i * N + j] = bw[j] / (bw[i] * (x[i] - x[j]));
        diag -= D[i * N + j];
      }
    D[i * N + i] = diag;
  }
}

// Exponential filter matrix V diag(exp(-alpha (i/(N-1))^(2 half_power))) V^-1
// with the Legendre Vandermonde matrix V(i,j) = P_j(x_i) at the LGL points and
// its numerical inverse (Spectral/Filtering.cpp:20-32, Spectral.cpp:498-523).
void exponential_filter_matrix(int N, double alpha, unsigned half_power,
                               std::vector<double>& F) {
  std::vector<double> x, w;
  lgl(N, x, w);
  std::vector<double> V((size_t)N * N), Vi((size_t)N * N, 0.0), A;
  for (int i = 0; i < N; ++i) {
    double pm2 = 1.0, pm1 = x[i];
    for (int j = 0; j < N; ++j) {
      double pj;
      if (j == 0) {
        pj = 1.0;
      } else if (j == 1) {
        pj = x[i];
      } else {
        pj = ((2.0 * j - 1.0) * x[i] * pm1 - (j - 1.0) * pm2) / j;
        pm2 = pm1;
        pm1 = pj;
      }
      V[(size_t)i * N + j] = pj;
    }
  }
  // Gauss-Jordan with partial pivoting
  A = V;
  for (int i = 0; i < N; ++i) Vi[(size_t)i * N + i] = 1.0;
  for (int c = 0; c < N; ++c) {
    int piv = c;
    for (int r = c + 1; r < N; ++r)
      if (std::abs(A[(size_t)r * N + c]) > std::abs(A[(size_t)piv * N + c])) piv = r;
    for (int k = 0; k < N; ++k) {
      std::swap(A[(size_t)c * N + k], A[(size_t)piv * N + k]);
      std::swap(Vi[(size_t)c * N + k], Vi[(size_t)piv * N + k]);
    }
    const double d = 1.0 / A[(size_t)c * N + c];
    for (int k = 0; k < N; ++k) {
      A[(size_t)c * N + k] *= d;
      Vi[(size_t)c * N + k] *= d;
    }
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const double fct = A[(size_t)r * N + c];
      if (fct == 0.0) continue;
      for (int k = 0; k < N; ++k) {
        A[(size_t)r * N + k] -= fct * A[(size_t)c * N + k];
        Vi[(size_t)r * N + k] -= fct * Vi[(size_t)c * N + k];
      }
    }
  }
  F.assign((size_t)N * N, 0.0);
  const double order = N - 1.0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int k = 0; k < N; ++k)
        s += V[(size_t)i * N + k] * std::exp(-alpha * std::pow(k / order, 2.0 * half_power)) *
             Vi[(size_t)k * N + j];
      F[(size_t)i * N + j] = s;
    }
}

// ---------------------------------------------------------------------------
// Adams-Bashforth coefficients (AdamsCoefficients.cpp:13-42, :75-117)
// ---------------------------------------------------------------------------
// orders 7 and 8: integrals over one step of the Lagrange basis through the k previous
// equally spaced times (the maximum order of the reference, AdamsBashforth.hpp:199)
const double kAbConst[9][8] = {
    {},
    {1.0},
    {-0.5, 1.5},
    {5.0 / 12.0, -4.0 / 3.0, 23.0 / 12.0},
    {-3.0 / 8.0, 37.0 / 24.0, -59.0 / 24.0, 55.0 / 24.0},
    {251.0 / 720.0, -637.0 / 360.0, 109.0 / 30.0, -1387.0 / 360.0, 1901.0 / 720.0},
    {-95.0 / 288.0, 959.0 / 480.0, -3649.0 / 720.0, 4991.0 / 720.0, -2641.0 / 480.0,
     4277.0 / 1440.0},
    {19087.0 / 60480.0, -5603.0 / 2520.0, 135713.0 / 20160.0, -10754.0 / 945.0,
     235183.0 / 20160.0, -18637.0 / 2520.0, 198721.0 / 60480.0},
    {-5257.0 / 17280.0, 32863.0 / 13440.0, -115747.0 / 13440.0, 2102243.0 / 120960.0,
     -296053.0 / 13440.0, 242653.0 / 13440.0, -1152169.0 / 120960.0, 16083.0 / 4480.0}};

std::vector<double> variable_coefficients(std::vector<double> ct, double step_start,
                                          double step_end) {
  for (auto& t : ct) t -= step_start;
  const size_t order = ct.size();
  std::vector<double> result;
  for (size_t j = 0; j < order; ++j) {
    std::vector<double> poly(order, 0.0);
    poly[0] = 1.0;
    for (size_t m = 0; m < order; ++m) {
      if (m == j) continue;
      const double denom = 1.0 / (ct[j] - ct[m]);
      for (size_t i = m < j ? m + 1 : m; i > 0; --i)
        poly[i] = (poly[i - 1] - poly[i] * ct[m]) * denom;
      poly[0] *= -ct[m] * denom;
    }
    for (size_t m = 0; m < order; ++m) poly[m] /= (double)(m + 1);
    const double dt = step_end - step_start;
    double val = 0.0;
    for (size_t m = order; m-- > 0;) val = val * dt + poly[m];
    result.push_back(dt * val);
  }
  return result;
}

// Butcher tableaus of the reference's RungeKutta steppers (RungeKutta.hpp:41-95):
// Rk3Owren.cpp:17-34, Rk3Kennedy.cpp:18-43, ClassicalRungeKutta4.cpp:24-49,
// DormandPrince5.cpp:21-50.  Only what the GTS path without error control uses:
// substep times c, substep coefficients A, result coefficients b; the number of
// substeps is b.size().
struct ButcherTableau {
  std::vector<double> substep_times;
  std::vector<std::vector<double>> substep_coefficients;
  std::vector<double> result_coefficients;
};

const ButcherTableau& butcher_tableau(int stepper) {
  static const ButcherTableau owren{
      {12.0 / 23.0, 4.0 / 5.0},
      {{12.0 / 23.0}, {-68.0 / 375.0, 368.0 / 375.0}},
      {31.0 / 144.0, 529.0 / 1152.0, 125.0 / 384.0}};
  static const ButcherTableau kennedy{
      {1767732205903.0 / 2027836641118.0, 3.0 / 5.0, 1.0},
      {{1767732205903.0 / 2027836641118.0},
       {5535828885825.0 / 10492691773637.0, 788022342437.0 / 10882634858940.0},
       {6485989280629.0 / 16251701735622.0, -4246266847089.0 / 9704473918619.0,
        10755448449292.0 / 10357097424841.0}},
      {1471266399579.0 / 7840856788654.0, -4482444167858.0 / 7529755066697.0,
       11266239266428.0 / 11593286722821.0, 1767732205903.0 / 4055673282236.0}};
  static const ButcherTableau rk4{
      {1.0 / 2.0, 1.0 / 2.0, 1.0, 3.0 / 4.0},
      {{1.0 / 2.0}, {0.0, 1.0 / 2.0}, {0.0, 0.0, 1.0},
       {5.0 / 32.0, 7.0 / 32.0, 13.0 / 32.0, -1.0 / 32.0}},
      {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0}};
  static const ButcherTableau dp5{
      {1.0 / 5.0, 3.0 / 10.0, 4.0 / 5.0, 8.0 / 9.0, 1.0, 1.0},
      {{1.0 / 5.0},
       {3.0 / 40.0, 9.0 / 40.0},
       {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0},
       {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0},
       {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0},
       {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}},
      {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
  switch (stepper) {
    case DGRHS_STEPPER_RK3_OWREN: return owren;
    case DGRHS_STEPPER_RK3_KENNEDY: return kennedy;
    case DGRHS_STEPPER_RK4: return rk4;
    default: return dp5;
  }
}

bool is_tableau_stepper(int stepper) {
  return stepper == DGRHS_STEPPER_RK3_OWREN || stepper == DGRHS_STEPPER_RK3_KENNEDY ||
         stepper == DGRHS_STEPPER_RK4 || stepper == DGRHS_STEPPER_DORMAND_PRINCE5;
}

// times in integer ticks of dt/tick_den
std::vector<double> ab_coefficients_ticks(const std::vector<long long>& ticks,
                                          long long start, long long end,
                                          long long tick_den, double dt) {
  const size_t order = ticks.size();
  bool uniform = ticks.back() == start;
  for (size_t i = 0; i + 1 < order; ++i)
    if (ticks[i + 1] - ticks[i] != end - start) uniform = false;
  auto frac = [&](long long t) { return (double)t / (double)tick_den; };
  if (uniform) {
    std::vector<double> c(order);
    for (size_t i = 0; i < order; ++i) c[i] = kAbConst[order][i] * (frac(end - start) * dt);
    return c;
  }
  std::vector<double> control{0.0};
  for (size_t i = 0; i + 1 < order; ++i)
    control.push_back(control.back() + frac(ticks[i + 1] - ticks[i]) * dt);
  return variable_coefficients(control, control.back() + frac(start - ticks.back()) * dt,
                               control.back() + frac(end - ticks.back()) * dt);
}

}  // namespace

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
namespace {

// host [E][ncomp][n] <-> device [E][ncomp][npad]
int upload(dgrhs_ctx* c, double* dst, const double* src, int ncomp) {
  CU(cudaMemcpy2DAsync(dst, (size_t)c->npad * 8, src, (size_t)c->n * 8, (size_t)c->n * 8,
                       (size_t)c->nelem * ncomp, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
int download(dgrhs_ctx* c, double* dst, const double* src, int ncomp) {
  CU(cudaMemcpy2DAsync(dst, (size_t)c->n * 8, src, (size_t)c->npad * 8, (size_t)c->n * 8,
                       (size_t)c->nelem * ncomp, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int rhs_range(dgrhs_ctx* c, double time, double* dt, int eb, int ee, bool volume_only,
              bool do_gauge, const dg::UpdateArgs& upd = dg::UpdateArgs{}) {
  c->pdl_volume = false;
  const DgNOps* ops = dgrhs_nops(c->N);
  if (!ops) return fail("unsupported number of grid points per dimension: %d", c->N);
  if (do_gauge && ops->gauge(c, time)) return 1;
  c->bjorhus_join_pending = false;
  if (!volume_only && ops->faces(c, eb, ee)) return 1;
  if (c->bjorhus_join_pending) {
    // The Bjorhus kernel (few long, latency-bound CTAs on the side stream) only feeds the
    // elements that own a Bjorhus face (the tail of the element order): the others go first
    c->bjorhus_join_pending = false;
    const int tb = c->bjorhus_tail_begin;
    int rc = ops->volume(c, dt, eb, tb, true, &upd);
    if (!rc && cudaStreamWaitEvent(c->stream, c->aux_join, 0) != cudaSuccess) rc = 1;
    c->pdl_volume = false;
    if (!rc) rc = ops->volume(c, dt, tb, ee, true, &upd);
    return rc;
  }
  if (ops->volume(c, dt, eb, ee, !volume_only, &upd)) return 1;
  if (c->mesh_v) {
    if (upd.u_new) return fail("internal error: fused update on a moving mesh");
    return ops->mesh_velocity_terms(c, dt, eb, ee);
  }
  return 0;
}

int lincomb(dgrhs_ctx* c, double* u, double a, const std::vector<double>& coef,
            const std::vector<const double*>& v) {
  if (coef.size() > 8) return fail("too many terms in linear combination");
  dg::LincombArgs p;
  p.u = u;
  p.a = a;
  p.nterms = (int)coef.size();
  for (size_t i = 0; i < coef.size(); ++i) {
    p.c[i] = coef[i];
    p.v[i] = v[i];
  }
  p.len2 = (long long)(c->state_len() / 2);
  const int blocks = (int)std::min<long long>((p.len2 + 255) / 256, 148LL * 16);
  dg::lincomb_kernel<<<blocks, 256, 0, c->stream>>>(p);
  ++g_launches;
  CU(cudaGetLastError());
  return 0;
}

}  // namespace

// constant_adams_moulton_coefficients (AdamsCoefficients.cpp:44-72), oldest value first, the
// value at the end of the step last
static const double kAmConst[9][8] = {
    {},
    {1.0},
    {0.5, 0.5},
    {-1.0 / 12.0, 2.0 / 3.0, 5.0 / 12.0},
    {1.0 / 24.0, -5.0 / 24.0, 19.0 / 24.0, 3.0 / 8.0},
    {-19.0 / 720.0, 53.0 / 360.0, -11.0 / 30.0, 323.0 / 360.0, 251.0 / 720.0},
    {3.0 / 160.0, -173.0 / 1440.0, 241.0 / 720.0, -133.0 / 240.0, 1427.0 / 1440.0, 95.0 / 288.0},
    {-863.0 / 60480.0, 263.0 / 2520.0, -6737.0 / 20160.0, 586.0 / 945.0, -15487.0 / 20160.0,
     2713.0 / 2520.0, 19087.0 / 60480.0},
    {275.0 / 24192.0, -11351.0 / 120960.0, 1537.0 / 4480.0, -88547.0 / 120960.0,
     123133.0 / 120960.0, -4511.0 / 4480.0, 139849.0 / 120960.0, 5257.0 / 17280.0}};

// adams_coefficients::coefficients (AdamsCoefficients.hpp:64-104): the constant-step tables
// when the control times are equally spaced by the step and the step starts (Adams-Bashforth)
// or ends (Adams-Moulton) at the last of them, else the integrals of the Lagrange polynomials
std::vector<double> dgrhs_internal_ab_coefficients_ticks(const std::vector<long long>& ticks,
                                                         long long start, long long end,
                                                         double tick_size) {
  const size_t order = ticks.size();
  bool uniform = true;
  for (size_t i = 0; i + 1 < order; ++i)
    if (ticks[i + 1] - ticks[i] != end - start) uniform = false;
  if (uniform && order <= 8 && ticks.back() == end && ticks.back() != start) {
    std::vector<double> c(order);
    for (size_t i = 0; i < order; ++i) c[i] = kAmConst[order][i] * ((double)(end - start) * tick_size);
    return c;
  }
  return ab_coefficients_ticks(ticks, start, end, 1, tick_size);
}

int dgrhs_internal_lincomb_range(dgrhs_ctx* c, double* u, double a,
                                 const std::vector<double>& coef,
                                 const std::vector<const double*>& v, size_t len) {
  if (coef.size() > 8) return fail("too many terms in linear combination");
  if (len == 0) return 0;
  dg::LincombArgs p;
  p.u = u;
  p.a = a;
  p.nterms = (int)coef.size();
  for (size_t i = 0; i < coef.size(); ++i) {
    p.c[i] = coef[i];
    p.v[i] = v[i];
  }
  p.len2 = (long long)(len / 2);
  const int blocks = (int)std::min<long long>((p.len2 + 255) / 256, 148LL * 16);
  dg::lincomb_kernel<<<blocks, 256, 0, c->stream>>>(p);
  ++g_launches;
  CU(cudaGetLastError());
  return 0;
}

int dgrhs_internal_upload(dgrhs_ctx* c, double* dst, const double* src, int ncomp) {
  return upload(c, dst, src, ncomp);
}

namespace {

int apply_filter(dgrhs_ctx* c) {
  if (!c->filterF) return 0;
  return dgrhs_nops(c->N)->filter(c);
}

int ensure_slots(dgrhs_ctx* c, int count) {
  while ((int)c->dt_slots.size() < count) {
    double* p = nullptr;
    if (dev_alloc(&p, c->state_len())) return 1;
    c->free_slots.push_back((int)c->dt_slots.size());
    c->dt_slots.push_back(p);
  }
  return 0;
}

void ab_clean(dgrhs_ctx* c, int order) {
  while ((int)c->history.size() >= order) {
    c->free_slots.push_back(c->history.front().slot);
    c->history.pop_front();
  }
}

int ab_update(dgrhs_ctx* c, int order, long long start, long long end) {
  std::vector<long long> ticks;
  std::vector<const double*> v;
  const size_t h = c->history.size();
  for (size_t i = h - order; i < h; ++i) {
    ticks.push_back(c->history[i].tick);
    v.push_back(c->dt_slots[c->history[i].slot]);
  }
  const auto coef = ab_coefficients_ticks(ticks, start, end, c->tick_den, c->dt);
  return lincomb(c, c->u, 1.0, coef, v);
}

// Prepare the stepper update that the volume kernel will fuse for the substep
// that begin_substep just set up.  Same coefficients and term order as the
// unfused ab_update / RK path, so the result is bit-identical.
int prepare_fused_update(dgrhs_ctx* c) {
  c->upd_active = false;
  if (!c->fuse_update || c->mesh_v) return 0;
  if (!c->u_alt && dev_alloc(&c->u_alt, c->state_len())) return 1;
  dg::UpdateArgs up{};
  up.u_new = c->u_alt;
  if (is_tableau_stepper(c->stepper)) return 0;  // separate update (up to 7 terms)
  if (c->stepper == DGRHS_STEPPER_ADAMS_BASHFORTH) {
    const SubstepOp& op = c->cur_op;
    if (op.kind != SubstepOp::kAbStep || op.order > 4) return 0;
    const size_t h = c->history.size();
    const size_t older = (size_t)op.order - 1;
    if (h < older) return fail("internal error: history too short for fused update");
    std::vector<long long> ticks;
    for (size_t i = h - older; i < h; ++i) ticks.push_back(c->history[i].tick);
    ticks.push_back(op.tick);
    const auto coef = ab_coefficients_ticks(ticks, op.tick, op.tick_end, c->tick_den, c->dt);
    up.a = 1.0;
    up.nterms = (int)older;
    for (size_t j = 0; j < older; ++j) {
      up.c[j] = coef[j];
      up.v[j] = c->dt_slots[c->history[h - older + j].slot];
    }
    up.c_new = coef.back();
  } else {
    const double dt = c->dt;
    if (c->rk_substep == 0) {
      up.a = 1.0;
      up.nterms = 0;
      up.c_new = dt;
    } else if (c->rk_substep == 1) {
      up.a = 0.25;
      up.nterms = 1;
      up.c[0] = 0.75;
      up.v[0] = c->u0;
      up.c_new = 0.25 * dt;
    } else {
      up.a = 2.0 / 3.0;
      up.nterms = 1;
      up.c[0] = 1.0 / 3.0;
      up.v[0] = c->u0;
      up.c_new = (2.0 / 3.0) * dt;
    }
  }
  c->pending_upd = up;
  c->upd_active = true;
  return 0;
}

}  // namespace


#define DG_FOR_EACH_N(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)
extern "C" {
#define X(NN) const DgNOps* dgrhs_internal_nops_##NN(void);
DG_FOR_EACH_N(X)
#undef X
}
const DgNOps* dgrhs_nops(int N) {
  switch (N) {
#define X(NN) \
  case NN:    \
    return dgrhs_internal_nops_##NN();
    DG_FOR_EACH_N(X)
#undef X
  }
  return nullptr;
}

// ---------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------
extern "C" {

// shared with operators.cu (not part of the public header)
void dgrhs_internal_set_error(const char* msg) { g_error = msg; }
void dgrhs_internal_count_launch(void) { ++g_launches; }
void dgrhs_internal_comm_destroy(void* comm);

const char* dgrhs_last_error(void) { return g_error.c_str(); }
int64_t dgrhs_kernel_launch_count(void) { return g_launches; }

int dgrhs_create(dgrhs_ctx** out, int system, int N, int nelem, int nghost, int device) {
  if (!out) return fail("null output pointer");
  if (system != DGRHS_SYSTEM_SCALAR_WAVE && system != DGRHS_SYSTEM_GH)
    return fail("unknown system %d", system);
  if (N < 2 || N > 12) return fail("n_points_1d must be in [2, 12], got %d", N);
  if (nelem < 1) return fail("n_elements must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device available: this library has no CPU fallback");
  CU(cudaSetDevice(device));
  dgrhs_ctx* c = new dgrhs_ctx();
  c->system = system;
  c->N = N;
  c->nelem = nelem;
  c->nghost = nghost;
  c->device = device;
  CU(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  c->C = system == DGRHS_SYSTEM_GH ? 50 : 5;
  c->S = system == DGRHS_SYSTEM_GH ? 3 : 1;
  c->HC = c->C + 3 + (system == DGRHS_SYSTEM_GH ? 2 : 1);
  c->n = N * N * N;
  c->f = N * N;
  c->npad = dg::padded_points(N);
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, hi));
    CU(cudaEventCreateWithFlags(&c->aux_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->aux_join, cudaEventDisableTiming));
  }
  if (dev_alloc(&c->u, c->state_len())) return 1;
  if (dev_alloc(&c->invjac, (size_t)nelem * 9 * c->npad)) return 1;
  if (dev_alloc(&c->stat, (size_t)nelem * c->S * c->npad)) return 1;
  if (dev_alloc(&c->corr, (size_t)nelem * 6 * c->C * c->f)) return 1;
  if (dev_alloc(&c->nbr, (size_t)nelem * 6)) return 1;
  if (dev_alloc(&c->D, (size_t)N * N)) return 1;
  if (nghost > 0) {
    if (dev_alloc(&c->halo_send, (size_t)nghost * c->HC * c->f)) return 1;
    if (dev_alloc(&c->halo_recv, (size_t)nghost * c->HC * c->f)) return 1;
    if (dev_alloc(&c->halo_map, (size_t)nghost * 2)) return 1;
  }
  std::vector<double> D;
  diff_matrix(N, D);
  CU(h2d_table(c->D, D.data(), D.size() * 8));
  std::vector<int32_t> nb((size_t)nelem * 6, -1);
  CU(h2d_table(c->nbr, nb.data(), nb.size() * 4));
  if (ensure_slots(c, 1)) return 1;
  c->dt_last = c->dt_slots[0];
  *out = c;
  return 0;
}

int dgrhs_destroy(dgrhs_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->nccl_comm) {
    cudaStreamSynchronize(c->comm_stream);
    dgrhs_internal_comm_destroy(c->nccl_comm);
    cudaStreamDestroy(c->comm_stream);
    cudaEventDestroy(c->ev_packed);
    cudaEventDestroy(c->ev_halo);
    cudaEventDestroy(c->ev_faces2);
    cudaEventDestroy(c->ev_faces1);
  }
  for (double* p : {c->u, c->invjac, c->coords, c->stat, c->corr, c->D, c->gH, c->gdH,
                    c->halo_send, c->halo_recv, c->u0, c->u_alt, c->ctxbuf, c->filterF,
                    c->mesh_v})
    if (p) cudaFree(p);
  for (double* p : c->dt_slots) cudaFree(p);
  dgrhs_internal_lts_free(c);
  if (c->nbr) cudaFree(c->nbr);
  if (c->nbr_face) cudaFree(c->nbr_face);
  if (c->violations) cudaFree(c->violations);
  if (c->bjorhus_faces) cudaFree(c->bjorhus_faces);
  if (c->mortar_faces) cudaFree(c->mortar_faces);
  if (c->mortar_table) cudaFree(c->mortar_table);
  if (c->mortar_P) cudaFree(c->mortar_P);
  for (void* q : {(void*)c->pm_faces, (void*)c->pm_ghost, (void*)c->pm_P, (void*)c->pm_R})
    if (q) cudaFree(q);
  if (c->pm_event) cudaEventDestroy(c->pm_event);
  if (c->mortar_R) cudaFree(c->mortar_R);
  if (c->halo_map) cudaFree(c->halo_map);
  if (c->aux_join) cudaEventDestroy(c->aux_join);
  if (c->aux_fork) cudaEventDestroy(c->aux_fork);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int dgrhs_set_geometry(dgrhs_ctx* c, const double* inv_jacobian, const double* coords,
                       const int32_t* neighbors) {
  CHECK_CTX(c);
  if (!inv_jacobian || !neighbors) return fail("inv_jacobian and neighbors are required");
  CU(cudaSetDevice(c->device));
  for (size_t i = 0; i < (size_t)c->nelem * 6; ++i) {
    const int v = neighbors[i];
    if (v == DGRHS_NEIGHBOR_HANGING) continue;  // non-conforming face: dgrhs_set_mortars
    if (v == DGRHS_NEIGHBOR_P_MORTAR) continue;  // neighbour with another N: dgrhs_set_p_mortars
    if (v == DGRHS_NEIGHBOR_BJORHUS || v == DGRHS_NEIGHBOR_BJORHUS_PHYSICAL) {
      if (c->system != DGRHS_SYSTEM_GH)
        return fail("ConstraintPreservingBjorhus is a GeneralizedHarmonic boundary condition");
      if (!coords) return fail("ConstraintPreservingBjorhus needs inertial coordinates");
      continue;
    }
    if (v >= c->nelem) return fail("neighbor index %d out of range", v);
    if (v <= -2 && -(v + 2) >= c->nghost) return fail("ghost face index out of range");
  }
  // conforming aligned interfaces: the neighbour's opposite face points back
  // (a table for rotated blocks is validated by dgrhs_set_neighbor_orientations,
  // which must then follow before the first right-hand side)
  c->aligned_table_ok = true;
  for (int e = 0; e < c->nelem && c->aligned_table_ok; ++e)
    for (int d = 0; d < 6; ++d) {
      const int v = neighbors[(size_t)e * 6 + d];
      if (v >= 0 && neighbors[(size_t)v * 6 + (d ^ 1)] != e) {
        c->aligned_table_ok = false;
        break;
      }
    }
  if (upload(c, c->invjac, inv_jacobian, 9)) return 1;
  if (coords) {
    if (!c->coords && dev_alloc(&c->coords, (size_t)c->nelem * 3 * c->npad)) return 1;
    if (upload(c, c->coords, coords, 3)) return 1;
  }
  CU(h2d_table(c->nbr, neighbors, (size_t)c->nelem * 6 * 4));
  c->nbr_host.assign(neighbors, neighbors + (size_t)c->nelem * 6);
  // a new neighbour table resets the orientations to "aligned" and drops the mortars
  if (c->nbr_face) cudaFree(c->nbr_face);
  c->nbr_face = nullptr;
  c->n_mortar_faces = 0;
  c->n_pmortar_faces = 0;
  // external faces with the Bjorhus boundary condition
  std::vector<int32_t> bj;
  for (int e = 0; e < c->nelem; ++e) {
    int dims_with_bjorhus = 0;
    for (int d = 0; d < 6; ++d)
      if (neighbors[(size_t)e * 6 + d] == DGRHS_NEIGHBOR_BJORHUS ||
          neighbors[(size_t)e * 6 + d] == DGRHS_NEIGHBOR_BJORHUS_PHYSICAL) {
        bj.insert(bj.end(),
                  {e, d, neighbors[(size_t)e * 6 + d] == DGRHS_NEIGHBOR_BJORHUS_PHYSICAL});
        dims_with_bjorhus |= 1 << (d >> 1);
      }
    // The reference applies the external faces of an element one after the other: a later
    // face projects the time derivative that earlier faces have already corrected on the
    // shared edge / corner points (BoundaryConditionsImpl.hpp:277-278, 636-660).  The
    // Bjorhus kernel evaluates every face from the uncorrected volume time derivative,
    // which is the same thing only if the faces share no points (opposite faces).
    if (dims_with_bjorhus & (dims_with_bjorhus - 1))
      return fail("element %d has ConstraintPreservingBjorhus faces in more than one dimension "
                  "(faces that share edge points): not supported", e);
  }
  if (c->bjorhus_faces) cudaFree(c->bjorhus_faces);
  c->bjorhus_faces = nullptr;
  c->n_bjorhus_faces = (int)(bj.size() / 3);
  c->bjorhus_tail_begin = -1;
  if (!bj.empty()) {
    int first = c->nelem;
    std::vector<char> has(c->nelem, 0);
    for (size_t k = 0; k < bj.size(); k += 3) {
      has[bj[k]] = 1;
      first = std::min(first, (int)bj[k]);
    }
    bool tail = first > 0;
    for (int e = first; e < c->nelem && tail; ++e) tail = has[e] != 0;
    if (tail) c->bjorhus_tail_begin = first;
  }
  if (!bj.empty()) {
    CU(cudaMalloc(&c->bjorhus_faces, bj.size() * 4));
    CU(h2d_table(c->bjorhus_faces, bj.data(), bj.size() * 4));
  }
  return 0;
}

int dgrhs_set_neighbor_orientations(dgrhs_ctx* c, const int32_t* neighbor_direction,
                                    const int32_t* face_permutation) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (c->nbr_host.empty()) return fail("call dgrhs_set_geometry first");
  std::vector<int32_t> packed((size_t)c->nelem * 6);
  auto apply = [&](int perm, int qa, int qb, int& na, int& nb) {
    na = (perm & 1) ? qb : qa;
    nb = (perm & 1) ? qa : qb;
    if (perm & 2) na = c->N - 1 - na;
    if (perm & 4) nb = c->N - 1 - nb;
  };
  for (int e = 0; e < c->nelem; ++e)
    for (int d = 0; d < 6; ++d) {
      const size_t k = (size_t)e * 6 + d;
      const int nd = neighbor_direction[k], perm = face_permutation[k];
      if (nd < 0 || nd > 5 || perm < 0 || perm > 7)
        return fail("bad orientation at element %d direction %d", e, d);
      packed[k] = nd | (perm << 3);
      const int v = c->nbr_host[k];
      if (v < 0) continue;  // external, ghost or hanging
      // the neighbour must point back at us with the inverse face map
      const size_t kn = (size_t)v * 6 + nd;
      if (c->nbr_host[kn] != e || neighbor_direction[kn] != d)
        return fail("orientations are not symmetric at element %d direction %d", e, d);
      for (int qb = 0; qb < c->N; ++qb)
        for (int qa = 0; qa < c->N; ++qa) {
          int na, nb, ba, bb;
          apply(perm, qa, qb, na, nb);
          apply(face_permutation[kn], na, nb, ba, bb);
          if (ba != qa || bb != qb)
            return fail("face permutations of element %d direction %d and its neighbour are "
                        "not inverse to each other", e, d);
        }
    }
  // the aligned default needs the opposite-face rule checked in set_geometry;
  // with explicit orientations that check is replaced by the one above
  if (!c->nbr_face && dev_alloc(&c->nbr_face, packed.size())) return 1;
  CU(h2d_table(c->nbr_face, packed.data(), packed.size() * 4));
  return 0;
}

// ---------------------------------------------------------------------------
// Non-conforming mortars: projection matrices (Spectral/Projection.cpp) and the
// mortar table.  The matrices are computed independently of the reference's
// closed forms: parent->child is barycentric interpolation at the child's LGL
// points mapped into the parent interval; child->parent is the exact L2
// projection V T V^-1 with T_kj = (2k+1)/2 int_child P_j(x_child) P_k(x) dx
// evaluated with an (N+1)-point LGL rule on the child interval (exact: the
// integrand has degree 2N-2).
// ---------------------------------------------------------------------------
namespace {
double legendre_p(int k, double x) {
  double pm2 = 1.0, pm1 = x;
  if (k == 0) return 1.0;
  if (k == 1) return x;
  double pk = 0.0;
  for (int j = 2; j <= k; ++j) {
    pk = ((2.0 * j - 1.0) * x * pm1 - (j - 1.0) * pm2) / j;
    pm2 = pm1;
    pm1 = pk;
  }
  return pk;
}

void invert(std::vector<double> A, int N, std::vector<double>& inv) {
  inv.assign((size_t)N * N, 0.0);
  for (int i = 0; i < N; ++i) inv[(size_t)i * N + i] = 1.0;
  for (int c = 0; c < N; ++c) {
    int piv = c;
    for (int r = c + 1; r < N; ++r)
      if (std::abs(A[(size_t)r * N + c]) > std::abs(A[(size_t)piv * N + c])) piv = r;
    for (int k = 0; k < N; ++k) {
      std::swap(A[(size_t)c * N + k], A[(size_t)piv * N + k]);
      std::swap(inv[(size_t)c * N + k], inv[(size_t)piv * N + k]);
    }
    const double d = 1.0 / A[(size_t)c * N + c];
    for (int k = 0; k < N; ++k) {
      A[(size_t)c * N + k] *= d;
      inv[(size_t)c * N + k] *= d;
    }
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const double fct = A[(size_t)r * N + c];
      if (fct == 0.0) continue;
      for (int k = 0; k < N; ++k) {
        A[(size_t)r * N + k] -= fct * A[(size_t)c * N + k];
        inv[(size_t)r * N + k] -= fct * inv[(size_t)c * N + k];
      }
    }
  }
}

// size: 0 Full, 1 LowerHalf, 2 UpperHalf; Np points on the parent (element face) mesh,
// Nc >= Np on the child (mortar) mesh.  Parent -> child [Nc][Np]: barycentric
// interpolation to the child's points mapped into the parent interval.  Child ->
// parent [Np][Nc]: the L2 projection, parent mode k = (2k+1)/2 * integral over the
// child's interval of (child function) * P_k, integrated exactly by Gauss-Lobatto
// quadrature with enough points for degree (Nc - 1) + (Np - 1).
void projection_matrix_meshes(int Np, int Nc, bool child_to_parent, int size,
                              std::vector<double>& M) {
  M.assign((size_t)Np * Nc, 0.0);
  if (size == 0 && Np == Nc) {
    for (int i = 0; i < Np; ++i) M[(size_t)i * Np + i] = 1.0;
    return;
  }
  std::vector<double> xp, wp, xc, wc;
  lgl(Np, xp, wp);
  lgl(Nc, xc, wc);
  // a child coordinate x sits at scale * x + shift in the parent interval
  const double scale = size == 0 ? 1.0 : 0.5;
  const double shift = size == 0 ? 0.0 : (size == 2 ? 0.5 : -0.5);
  if (!child_to_parent) {
    std::vector<double> bw(Np, 1.0);
    for (int j = 1; j < Np; ++j)
      for (int k = 0; k < j; ++k) {
        bw[k] *= xp[k] - xp[j];
        bw[j] *= xp[j] - xp[k];
      }
    for (int j = 0; j < Np; ++j) bw[j] = 1.0 / bw[j];
    for (int k = 0; k < Nc; ++k) {
      const double t = scale * xc[k] + shift;
      int match = -1;
      for (int j = 0; j < Np; ++j)
        if (std::abs(t - xp[j]) < 1e-14) match = j;
      if (match >= 0) {
        M[(size_t)k * Np + match] = 1.0;
        continue;
      }
      double sum = 0.0;
      for (int j = 0; j < Np; ++j) {
        M[(size_t)k * Np + j] = bw[j] / (t - xp[j]);
        sum += M[(size_t)k * Np + j];
      }
      for (int j = 0; j < Np; ++j) M[(size_t)k * Np + j] /= sum;
    }
    return;
  }
  // nodal (child) -> modal (child) -> modal (parent) -> nodal (parent)
  std::vector<double> Vc((size_t)Nc * Nc), Vci, xq, wq;
  for (int i = 0; i < Nc; ++i)
    for (int j = 0; j < Nc; ++j) Vc[(size_t)i * Nc + j] = legendre_p(j, xc[i]);
  invert(Vc, Nc, Vci);
  const int nq = (Np + Nc + 1) / 2 + 1;
  lgl(nq, xq, wq);
  std::vector<double> T((size_t)Np * Nc, 0.0);  // parent mode k from child mode j
  for (int k = 0; k < Np; ++k)
    for (int j = 0; j < Nc; ++j) {
      double sum = 0.0;
      for (int q = 0; q < nq; ++q)
        sum += scale * wq[q] * legendre_p(j, xq[q]) * legendre_p(k, scale * xq[q] + shift);
      T[(size_t)k * Nc + j] = 0.5 * (2.0 * k + 1.0) * sum;
    }
  std::vector<double> TV((size_t)Np * Nc, 0.0);  // parent modes from child nodal values
  for (int k = 0; k < Np; ++k)
    for (int j = 0; j < Nc; ++j) {
      double sum = 0.0;
      for (int m = 0; m < Nc; ++m) sum += T[(size_t)k * Nc + m] * Vci[(size_t)m * Nc + j];
      TV[(size_t)k * Nc + j] = sum;
    }
  for (int i = 0; i < Np; ++i)
    for (int j = 0; j < Nc; ++j) {
      double sum = 0.0;
      for (int k = 0; k < Np; ++k) sum += legendre_p(k, xp[i]) * TV[(size_t)k * Nc + j];
      M[(size_t)i * Nc + j] = sum;
    }
}
// size: 0 Full, 1 LowerHalf, 2 UpperHalf; same number of points on both meshes
void projection_matrix(int N, bool child_to_parent, int size, std::vector<double>& M) {
  M.assign((size_t)N * N, 0.0);
  if (size == 0) {
    for (int i = 0; i < N; ++i) M[(size_t)i * N + i] = 1.0;
    return;
  }
  std::vector<double> x, w;
  lgl(N, x, w);
  const double shift = size == 2 ? 1.0 : -1.0;  // child point xc sits at (xc + shift) / 2
  if (!child_to_parent) {
    std::vector<double> bw(N, 1.0);
    for (int j = 1; j < N; ++j)
      for (int k = 0; k < j; ++k) {
        bw[k] *= x[k] - x[j];
        bw[j] *= x[j] - x[k];
      }
    for (int j = 0; j < N; ++j) bw[j] = 1.0 / bw[j];
    for (int k = 0; k < N; ++k) {
      const double t = 0.5 * (x[k] + shift);
      int match = -1;
      for (int j = 0; j < N; ++j)
        if (std::abs(t - x[j]) < 1e-14) match = j;
      if (match >= 0) {
        M[(size_t)k * N + match] = 1.0;
        continue;
      }
      double sum = 0.0;
      for (int j = 0; j < N; ++j) {
        M[(size_t)k * N + j] = bw[j] / (t - x[j]);
        sum += M[(size_t)k * N + j];
      }
      for (int j = 0; j < N; ++j) M[(size_t)k * N + j] /= sum;
    }
    return;
  }
  std::vector<double> V((size_t)N * N), Vi, xq, wq;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[(size_t)i * N + j] = legendre_p(j, x[i]);
  invert(V, N, Vi);
  lgl(N + 1, xq, wq);
  std::vector<double> T((size_t)N * N, 0.0);
  for (int k = 0; k < N; ++k)
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int q = 0; q <= N; ++q)  // child coordinate xq, parent coordinate (xq + shift)/2
        s += 0.5 * wq[q] * legendre_p(j, xq[q]) * legendre_p(k, 0.5 * (xq[q] + shift));
      T[(size_t)k * N + j] = 0.5 * (2.0 * k + 1.0) * s;
    }
  std::vector<double> VT((size_t)N * N, 0.0);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int k = 0; k < N; ++k) s += V[(size_t)i * N + k] * T[(size_t)k * N + j];
      VT[(size_t)i * N + j] = s;
    }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int k = 0; k < N; ++k) s += VT[(size_t)i * N + k] * Vi[(size_t)k * N + j];
      M[(size_t)i * N + j] = s;
    }
}
}  // namespace

int dgrhs_projection_matrix(int N, int child_to_parent, int size, double* matrix) {
  if (N < 2 || N > 12) return fail("n_points_1d must be in [2, 12]");
  if (size < 0 || size > 2) return fail("mortar size must be 0 (Full), 1 (LowerHalf) or 2 (UpperHalf)");
  std::vector<double> M;
  projection_matrix(N, child_to_parent != 0, size, M);
  std::memcpy(matrix, M.data(), M.size() * 8);
  return 0;
}

int dgrhs_projection_matrix_meshes(int n_parent, int n_child, int child_to_parent, int size,
                                   double* matrix) {
  if (n_parent < 2 || n_child > 12 || n_child < n_parent)
    return fail("need 2 <= n_parent <= n_child <= 12 (the mortar mesh is the finer one)");
  if (size < 0 || size > 2) return fail("mortar size must be 0 (Full), 1 (LowerHalf) or 2 (UpperHalf)");
  std::vector<double> M;
  if (n_parent == n_child)  // the matrices the mortar kernel uses
    projection_matrix(n_parent, child_to_parent != 0, size, M);
  else
    projection_matrix_meshes(n_parent, n_child, child_to_parent != 0, size, M);
  std::memcpy(matrix, M.data(), M.size() * 8);
  return 0;
}

// ---- p-refinement: faces between contexts with different N ------------------------
int dgrhs_set_p_mortars(dgrhs_ctx* c, int n_faces, const int32_t* table) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (c->nbr_host.empty()) return fail("call dgrhs_set_geometry first");
  if (n_faces < 0 || (n_faces > 0 && !table)) return fail("bad p-mortar table");
  size_t marked = 0;
  for (int v : c->nbr_host) marked += v == DGRHS_NEIGHBOR_P_MORTAR;
  if (marked != (size_t)n_faces)
    return fail("%zu faces are marked DGRHS_NEIGHBOR_P_MORTAR but the table lists %d", marked,
                n_faces);
  std::vector<char> seen((size_t)c->nelem * 6, 0);
  for (int i = 0; i < n_faces; ++i) {
    const int32_t* r = table + 4 * (size_t)i;
    if (r[0] < 0 || r[0] >= c->nelem || r[1] < 0 || r[1] > 5) return fail("p-mortar %d: bad face", i);
    if (c->nbr_host[(size_t)r[0] * 6 + r[1]] != DGRHS_NEIGHBOR_P_MORTAR || seen[(size_t)r[0] * 6 + r[1]]++)
      return fail("p-mortar %d: the face must be marked DGRHS_NEIGHBOR_P_MORTAR (once)", i);
    if (r[2] < 2 || r[2] > 12 || r[2] == c->N)
      return fail("p-mortar %d: the neighbour needs 2..12 grid points per dimension, not %d "
                  "(equal N: a conforming face of one context)", i, r[2]);
    if ((r[3] & 7) > 5 || (r[3] >> 3) < 0 || (r[3] >> 3) > 7) return fail("p-mortar %d: bad orientation", i);
  }
  for (void** q : {(void**)&c->pm_faces, (void**)&c->pm_ghost})
    if (*q) {
      cudaFree(*q);
      *q = nullptr;
    }
  c->n_pmortar_faces = n_faces;
  if (n_faces == 0) return 0;
  CU(cudaMalloc(&c->pm_faces, (size_t)n_faces * 16));
  CU(h2d_table(c->pm_faces, table, (size_t)n_faces * 16));
  if (dev_alloc(&c->pm_ghost, (size_t)n_faces * c->HC * 144)) return 1;
  // mortar mesh = the larger extents (MortarHelpers.cpp:22-49): interpolation of the side
  // with fewer points up to it, L2 projection back (Projection.cpp:57-362)
  std::vector<double> P(13 * 144, 0.0), R(13 * 144, 0.0), M;
  for (int nb = 2; nb <= 12; ++nb) {
    if (nb == c->N) continue;
    const int lo = std::min(nb, c->N), hi = std::max(nb, c->N);
    projection_matrix_meshes(lo, hi, false, 0, M);
    std::copy(M.begin(), M.end(), P.begin() + (size_t)nb * 144);
    projection_matrix_meshes(lo, hi, true, 0, M);
    std::copy(M.begin(), M.end(), R.begin() + (size_t)nb * 144);
  }
  if (!c->pm_P && dev_alloc(&c->pm_P, P.size())) return 1;
  if (!c->pm_R && dev_alloc(&c->pm_R, R.size())) return 1;
  CU(h2d_table(c->pm_P, P.data(), P.size() * 8));
  CU(h2d_table(c->pm_R, R.data(), R.size() * 8));
  if (!c->pm_event) CU(cudaEventCreateWithFlags(&c->pm_event, cudaEventDisableTiming));
  return 0;
}

int dgrhs_p_mortar_transfer(dgrhs_ctx* src, dgrhs_ctx* dst, int n, const int32_t* src_slots,
                            const int32_t* dst_faces) {
  CHECK_CTX(src);
  CHECK_CTX(dst);
  if (src->device != dst->device) return fail("p-mortar transfer: contexts on different devices");
  if (src->system != dst->system) return fail("p-mortar transfer: different evolution systems");
  if (n < 0 || (n > 0 && (!src_slots || !dst_faces))) return fail("bad arguments");
  if (n == 0) return 0;
  if (!src->pm_event) CU(cudaEventCreateWithFlags(&src->pm_event, cudaEventDisableTiming));
  CU(cudaSetDevice(dst->device));
  // after the source's dgrhs_pack_halo, before the destination's next right-hand side
  CU(cudaEventRecord(src->pm_event, src->stream));
  CU(cudaStreamWaitEvent(dst->stream, src->pm_event, 0));
  const size_t per_src = (size_t)src->HC * src->f;
  for (int i = 0; i < n; ++i) {
    if (src_slots[i] < 0 || src_slots[i] >= src->n_send) return fail("p-mortar transfer: bad halo slot");
    if (dst_faces[i] < 0 || dst_faces[i] >= dst->n_pmortar_faces) return fail("p-mortar transfer: bad face");
    CU(cudaMemcpyAsync(dst->pm_ghost + (size_t)dst_faces[i] * dst->HC * 144,
                       src->halo_send + (size_t)src_slots[i] * per_src, per_src * 8,
                       cudaMemcpyDeviceToDevice, dst->stream));
  }
  // the source must not pack again before the copies have read its send buffer
  CU(cudaEventRecord(dst->pm_event, dst->stream));
  CU(cudaStreamWaitEvent(src->stream, dst->pm_event, 0));
  return 0;
}

int dgrhs_set_mortars(dgrhs_ctx* c, int n_mortars, const int32_t* mortars) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (c->nbr_host.empty()) return fail("call dgrhs_set_geometry first");
  if (n_mortars < 0 || (n_mortars > 0 && !mortars)) return fail("bad mortar table");
  // a side may live on another rank: element = -(slot + 2) names the ghost slot
  // that receives its face.  Groups (the mortars of one coarse face) keep the
  // caller's order inside; groups without any remote side come first (they run
  // before the halo arrives), the others after them.
  auto is_ghost = [](int e) { return e <= -2; };
  std::vector<int> order(n_mortars);
  for (int m = 0; m < n_mortars; ++m) order[m] = m;
  // group key: a local coarse face (element, direction); a remote coarse side is
  // its own group (one ghost slot per mortar)
  auto key = [&](int m) {
    const int32_t* r = mortars + 6 * (size_t)m;
    return is_ghost(r[0]) ? std::pair<long long, int>((long long)c->nelem + (-(r[0] + 2)), r[1])
                          : std::pair<long long, int>(r[0], r[1]);
  };
  std::vector<char> coarse_remote(n_mortars, 0);
  {
    std::vector<std::pair<std::pair<long long, int>, int>> keyed;
    for (int m = 0; m < n_mortars; ++m) keyed.push_back({key(m), m});
    std::stable_sort(keyed.begin(), keyed.end(),
                     [](const auto& x, const auto& y) { return x.first < y.first; });
    // a group is "remote" if any of its sides is a ghost
    size_t i = 0;
    std::vector<std::pair<int, std::vector<int>>> groups;  // (remote flag, members)
    while (i < keyed.size()) {
      size_t j = i;
      int remote = 0;
      std::vector<int> members;
      while (j < keyed.size() && keyed[j].first == keyed[i].first) {
        const int32_t* r = mortars + 6 * (size_t)keyed[j].second;
        remote |= is_ghost(r[0]) || is_ghost(r[2]);
        members.push_back(keyed[j].second);
        ++j;
      }
      groups.push_back({remote, members});
      i = j;
    }
    std::stable_sort(groups.begin(), groups.end(),
                     [](const auto& x, const auto& y) { return x.first < y.first; });
    order.clear();
    c->n_mortar_faces_local = 0;
    for (const auto& g : groups) {
      if (!g.first) ++c->n_mortar_faces_local;
      for (int m : g.second) order.push_back(m);
    }
  }
  std::vector<int32_t> faces, table;
  std::vector<char> fine_seen((size_t)c->nelem * 6, 0);
  size_t local_coarse = 0, local_fine = 0;
  std::pair<long long, int> last_key{-1, -1};
  for (int k = 0; k < n_mortars; ++k) {
    const int32_t* m = mortars + 6 * (size_t)order[k];
    // m[3] = fine direction | perm << 3 (perm: the face permutation of non-aligned
    // blocks, bits as in dgrhs_set_neighbor_orientations, taking a mortar point in the
    // coarse element's face frame to the fine element's face point)
    const int ec = m[0], dc = m[1], ef = m[2], df = m[3] & 7, permf = m[3] >> 3, sa = m[4],
              sb = m[5];
    if (ec >= c->nelem || ef >= c->nelem || ec == -1 || ef == -1 || dc < 0 || dc > 5 ||
        m[3] < 0 || df > 5 || permf > 7)
      return fail("mortar %d: element, direction or face permutation out of range", order[k]);
    if ((is_ghost(ec) && -(ec + 2) >= c->nghost) || (is_ghost(ef) && -(ef + 2) >= c->nghost))
      return fail("mortar %d: ghost face index out of range", order[k]);
    if (is_ghost(ec) && is_ghost(ef)) return fail("mortar %d: both sides are remote", order[k]);
    if (sa < 0 || sa > 2 || sb < 0 || sb > 2 || (sa == 0 && sb == 0))
      return fail("mortar %d: bad mortar size (%d, %d)", order[k], sa, sb);
    if ((ec >= 0 && c->nbr_host[(size_t)ec * 6 + dc] != DGRHS_NEIGHBOR_HANGING) ||
        (ef >= 0 && c->nbr_host[(size_t)ef * 6 + df] != DGRHS_NEIGHBOR_HANGING))
      return fail("mortar %d: both faces must be marked DGRHS_NEIGHBOR_HANGING in the "
                  "neighbor table", order[k]);
    if (ef >= 0) {
      if (fine_seen[(size_t)ef * 6 + df]++)
        return fail("mortar %d: fine face listed twice", order[k]);
      ++local_fine;
    }
    const auto kk = key(order[k]);
    if (kk != last_key) {
      faces.insert(faces.end(), {ec, dc, k, 0});
      last_key = kk;
      if (ec >= 0) ++local_coarse;
    }
    ++faces[faces.size() - 1];
    table.insert(table.end(), {ef, m[3], sa, sb});
  }
  // every hanging face must be covered, or the volume kernel would add stale data
  size_t hanging = 0;
  for (int v : c->nbr_host) hanging += v == DGRHS_NEIGHBOR_HANGING;
  if (hanging != local_coarse + local_fine)
    return fail("%zu faces are marked hanging but the mortar table covers %zu", hanging,
                local_coarse + local_fine);
  c->mortar_faces_host = faces;
  c->mortar_table_host = table;
  if (c->mortar_faces) cudaFree(c->mortar_faces);
  if (c->mortar_table) cudaFree(c->mortar_table);
  c->mortar_faces = c->mortar_table = nullptr;
  c->n_mortar_faces = (int)(faces.size() / 4);
  if (n_mortars == 0) return 0;
  CU(cudaMalloc(&c->mortar_faces, faces.size() * 4));
  CU(cudaMalloc(&c->mortar_table, table.size() * 4));
  CU(h2d_table(c->mortar_faces, faces.data(), faces.size() * 4));
  CU(h2d_table(c->mortar_table, table.data(), table.size() * 4));
  const int N = c->N;
  std::vector<double> P, R, M;
  for (int size = 0; size < 3; ++size) {
    projection_matrix(N, false, size, M);
    P.insert(P.end(), M.begin(), M.end());
    projection_matrix(N, true, size, M);
    R.insert(R.end(), M.begin(), M.end());
  }
  if (!c->mortar_P && dev_alloc(&c->mortar_P, P.size())) return 1;
  if (!c->mortar_R && dev_alloc(&c->mortar_R, R.size())) return 1;
  CU(h2d_table(c->mortar_P, P.data(), P.size() * 8));
  CU(h2d_table(c->mortar_R, R.data(), R.size() * 8));
  return 0;
}

int dgrhs_set_static_fields(dgrhs_ctx* c, const double* fields, int ncomp) {
  CHECK_CTX(c);
  if (ncomp != c->S) return fail("expected %d static components, got %d", c->S, ncomp);
  CU(cudaSetDevice(c->device));
  return upload(c, c->stat, fields, ncomp);
}

int dgrhs_set_mesh_velocity(dgrhs_ctx* c, const double* mesh_velocity) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (c->in_substep) return fail("dgrhs_set_mesh_velocity inside a substep");
  if (!mesh_velocity) {
    CU(cudaStreamSynchronize(c->stream));
    if (c->mesh_v) cudaFree(c->mesh_v);
    c->mesh_v = nullptr;
    return 0;
  }
  if (c->n_bjorhus_faces > 0 || c->n_mortar_faces > 0 || c->n_pmortar_faces > 0)
    return fail("moving mesh: Bjorhus faces and non-conforming mortars are not supported");
  if (!c->mesh_v && dev_alloc(&c->mesh_v, (size_t)c->nelem * 3 * c->npad)) return 1;
  return upload(c, c->mesh_v, mesh_velocity, 3);
}

int dgrhs_set_gauge(dgrhs_ctx* c, int gauge, const double* params, int nparams) {
  CHECK_CTX(c);
  if (c->system != DGRHS_SYSTEM_GH) return fail("gauge conditions only apply to GH");
  if (gauge < 0 || gauge > 3) return fail("unknown gauge %d", gauge);
  if (nparams > 8) return fail("too many gauge parameters");
  CU(cudaSetDevice(c->device));
  c->gauge = gauge;
  for (int i = 0; i < nparams; ++i) c->gauge_params[i] = params[i];
  if (gauge == DGRHS_GAUGE_DAMPED_HARMONIC && nparams != 7)
    return fail("DampedHarmonic needs {width, amp_L1, amp_L2, amp_S, exp_L1, exp_L2, exp_S}");
  if ((gauge == DGRHS_GAUGE_FIELDS || gauge == DGRHS_GAUGE_ANALYTIC_GAUGE_WAVE) && !c->gH) {
    if (dev_alloc(&c->gH, (size_t)c->nelem * 4 * c->npad)) return 1;
    if (dev_alloc(&c->gdH, (size_t)c->nelem * 16 * c->npad)) return 1;
  }
  if (gauge == DGRHS_GAUGE_ANALYTIC_GAUGE_WAVE && nparams != 2)
    return fail("AnalyticChristoffel(GaugeWave) needs {amplitude, wavelength}");
  return 0;
}

int dgrhs_set_gauge_analytic_christoffel(dgrhs_ctx* c, const double* u_analytic) {
  CHECK_CTX(c);
  if (c->system != DGRHS_SYSTEM_GH) return fail("gauge conditions only apply to GH");
  CU(cudaSetDevice(c->device));
  if (dgrhs_set_gauge(c, DGRHS_GAUGE_FIELDS, nullptr, 0)) return 1;
  double* tmp = nullptr;
  if (dev_alloc(&tmp, c->state_len())) return 1;
  if (upload(c, tmp, u_analytic, c->C)) return 1;
  const int rc = dgrhs_nops(c->N)->gauge_from_state(c, tmp);
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(tmp);
  return rc;
}

int dgrhs_set_gauge_fields(dgrhs_ctx* c, const double* H, const double* dH) {
  CHECK_CTX(c);
  if (c->gauge != DGRHS_GAUGE_FIELDS) return fail("gauge is not DGRHS_GAUGE_FIELDS");
  CU(cudaSetDevice(c->device));
  if (upload(c, c->gH, H, 4)) return 1;
  return upload(c, c->gdH, dH, 16);
}

int dgrhs_set_state(dgrhs_ctx* c, const double* u) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  return upload(c, c->u, u, c->C);
}
int dgrhs_get_state(dgrhs_ctx* c, double* u) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  return download(c, u, c->u, c->C);
}
// stream-ordered forms: the copy is queued behind the context's earlier work and the
// call returns at once (page-locked host memory needed for a truly asynchronous copy)
int dgrhs_set_state_async(dgrhs_ctx* c, const double* u) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy2DAsync(c->u, (size_t)c->npad * 8, u, (size_t)c->n * 8, (size_t)c->n * 8,
                       (size_t)c->nelem * c->C, cudaMemcpyHostToDevice, c->stream));
  return 0;
}
int dgrhs_get_state_async(dgrhs_ctx* c, double* u) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy2DAsync(u, (size_t)c->n * 8, c->u, (size_t)c->npad * 8, (size_t)c->n * 8,
                       (size_t)c->nelem * c->C, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}
int dgrhs_get_time_derivative(dgrhs_ctx* c, double* dt_u) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  return download(c, dt_u, c->dt_last, c->C);
}

int dgrhs_compute_time_derivative(dgrhs_ctx* c, double time, int volume_only) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (c->n_send > 0 && !volume_only)
    return fail("context exchanges faces with other ranks: use pack_halo + "
                "compute_time_derivative_range");
  ++c->rhs_evals;
  // inside a substep (begin_substep ... end_substep) the stepper update may be fused into
  // the volume kernel; a volume-only evaluation is a diagnostic and never is
  if (c->in_substep && c->upd_active && volume_only)
    return fail("volume_only evaluation inside a substep with the fused update");
  return rhs_range(c, time, c->dt_last, 0, c->nelem, volume_only != 0, true,
                   (c->in_substep && c->upd_active) ? c->pending_upd : dg::UpdateArgs{});
}

int dgrhs_set_interior_count(dgrhs_ctx* c, int n_interior) {
  CHECK_CTX(c);
  if (n_interior < 0 || n_interior > c->nelem) return fail("bad interior count");
  c->n_interior = n_interior;
  return 0;
}

int dgrhs_set_halo_map(dgrhs_ctx* c, const int32_t* map, int n_send) {
  CHECK_CTX(c);
  if (n_send < 0 || n_send > c->nghost) return fail("n_send must be in [0, n_ghost_faces]");
  c->n_send = n_send;
  if (n_send == 0) return 0;
  CU(cudaSetDevice(c->device));
  for (int i = 0; i < n_send; ++i)
    if (map[2 * i] < 0 || map[2 * i] >= c->nelem || map[2 * i + 1] < 0 || map[2 * i + 1] > 5)
      return fail("bad halo map entry %d", i);
  CU(h2d_table(c->halo_map, map, (size_t)n_send * 8));
  return 0;
}

int dgrhs_set_boundary_ghost_data(dgrhs_ctx* c, int slot_begin, int n_slots,
                                  const double* data) {
  CHECK_CTX(c);
  if (slot_begin < 0 || n_slots < 0 || slot_begin + n_slots > c->nghost)
    return fail("ghost slots [%d, %d) out of range", slot_begin, slot_begin + n_slots);
  if (n_slots == 0) return 0;
  CU(cudaSetDevice(c->device));
  const size_t per = (size_t)c->HC * c->f;
  // ordered after the kernels already queued on the context's stream
  CU(cudaMemcpyAsync(c->halo_recv + (size_t)slot_begin * per, data, (size_t)n_slots * per * 8,
                     cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int dgrhs_pack_halo(dgrhs_ctx* c) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  return dgrhs_nops(c->N)->pack(c);
}

int dgrhs_compute_time_derivative_range(dgrhs_ctx* c, double time, int eb, int ee) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (eb < 0 || ee > c->nelem || eb > ee) return fail("bad element range");
  if (eb == ee) return 0;  // an empty range (e.g. no interior elements) is a no-op
  // the first non-empty range of an RHS evaluation counts it and runs the gauge
  // kernels; the evaluation is complete once every element has been covered
  const bool first = c->range_covered == 0;
  if (first) ++c->rhs_evals;
  c->range_covered += ee - eb;
  if (c->range_covered >= c->nelem) c->range_covered = 0;
  return rhs_range(c, time, c->dt_last, eb, ee, false, first,
                   (c->in_substep && c->upd_active) ? c->pending_upd : dg::UpdateArgs{});
}

void* dgrhs_halo_send_ptr(dgrhs_ctx* c) { return c ? c->halo_send : nullptr; }
void* dgrhs_halo_recv_ptr(dgrhs_ctx* c) { return c ? c->halo_recv : nullptr; }
int dgrhs_halo_comps(dgrhs_ctx* c) { return c ? c->HC : 0; }

// ---- time stepping --------------------------------------------------------

int dgrhs_set_stepper(dgrhs_ctx* c, int stepper, int order, double t0, double dt) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (stepper == DGRHS_STEPPER_ADAMS_BASHFORTH) {
    if (order < 1 || order > 8) return fail("Adams-Bashforth order must be in [1, 8]");
  } else if (stepper == DGRHS_STEPPER_RK3_HESTHAVEN) {
    order = 3;
  } else if (is_tableau_stepper(stepper)) {
    order = (int)butcher_tableau(stepper).result_coefficients.size();  // substeps
  } else {
    return fail("unknown stepper %d", stepper);
  }
  c->stepper = stepper;
  c->order = order;
  c->t0 = t0;
  c->dt = dt;
  c->step_index = 0;
  c->steps_per_slab = 0;
  c->rk_substep = 0;
  c->in_substep = false;
  c->history.clear();
  c->pending.clear();
  c->free_slots.clear();
  for (int i = 0; i < (int)c->dt_slots.size(); ++i) c->free_slots.push_back(i);
  if (!c->u0 && dev_alloc(&c->u0, c->state_len())) return 1;
  if (stepper == DGRHS_STEPPER_ADAMS_BASHFORTH) {
    c->tick_den = order;
    if (ensure_slots(c, order)) return 1;
    // forward self-start program (SelfStartActions.hpp; see oracle/oracle.py
    // Evolution._self_start for the trace of the action list)
    if (order > 1) {
      for (int o = 1; o < order; ++o) {
        c->pending.push_back({SubstepOp::kRestoreU0, o, 0, 0});
        for (int s = 0; s <= o; ++s) {
          if (s == o)
            c->pending.push_back({SubstepOp::kAbEvalOnly, o, (long long)s, 0});
          else
            c->pending.push_back({SubstepOp::kAbStep, o, (long long)s, (long long)s + 1});
        }
      }
      c->pending.push_back({SubstepOp::kRestoreU0, order, 0, 0});
      CU(cudaMemcpyAsync(c->u0, c->u, c->state_len() * 8, cudaMemcpyDeviceToDevice,
                         c->stream));
    }
  } else if (is_tableau_stepper(stepper)) {
    c->tick_den = 1;
    if (ensure_slots(c, order)) return 1;  // one derivative slot per substep
  } else {
    c->tick_den = 2;
    if (ensure_slots(c, 1)) return 1;
  }
  return 0;
}

// Time of step boundary `tick` (in units of dt / tick_den) the way the reference forms it:
// a Time is a slab plus an exact rational fraction f of it, value (1 - f) start + f end
// (Time.cpp:114-117); the slab that follows [a, b] is [b, b + (b - a)] (Slab.hpp advance).
// Without dgrhs_set_slab: t0 + (tick / tick_den) dt.
static double time_at_tick(const dgrhs_ctx* c, long long tick) {
  if (c->steps_per_slab == 0) return c->t0 + ((double)tick / (double)c->tick_den) * c->dt;
  const long long den = (long long)c->steps_per_slab * c->tick_den;
  long long slab = tick / den, num = tick % den;
  double a = c->slab_start, b = c->slab_end;
  for (long long s = 0; s < slab; ++s) {
    const double next = b + (b - a);
    a = b;
    b = next;
  }
  return ((double)(den - num) / (double)den) * a + ((double)num / (double)den) * b;
}

int dgrhs_set_slab(dgrhs_ctx* c, double slab_start, double slab_end, int steps_per_slab) {
  CHECK_CTX(c);
  if (c->dt == 0.0) return fail("set_stepper has not been called");
  if (c->in_substep || c->step_index != 0) return fail("set_slab must follow set_stepper directly");
  if (!(slab_start < slab_end) || steps_per_slab < 1) return fail("bad slab");
  c->slab_start = slab_start;
  c->slab_end = slab_end;
  c->steps_per_slab = steps_per_slab;
  c->t0 = slab_start;
  // TimeDelta::value (Time.cpp:127-129): slab duration times the fraction's double value
  c->dt = (slab_end - slab_start) * (1.0 / (double)steps_per_slab);
  return 0;
}

int dgrhs_self_start_substeps_left(dgrhs_ctx* c, int* n) {
  CHECK_CTX(c);
  int k = 0;
  for (const SubstepOp& op : c->pending) k += op.kind != SubstepOp::kRestoreU0;
  *n = k;
  return 0;
}

int dgrhs_begin_substep(dgrhs_ctx* c, double* time) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (c->in_substep) return fail("begin_substep called twice");
  if (c->dt == 0.0) return fail("set_stepper has not been called");
  c->range_covered = 0;
  if (c->stepper == DGRHS_STEPPER_ADAMS_BASHFORTH) {
    while (!c->pending.empty() && c->pending.front().kind == SubstepOp::kRestoreU0) {
      CU(cudaMemcpyAsync(c->u, c->u0, c->state_len() * 8, cudaMemcpyDeviceToDevice,
                         c->stream));
      c->pending.pop_front();
    }
    if (!c->pending.empty()) {
      c->cur_op = c->pending.front();
      c->pending.pop_front();
    } else {
      const long long t = c->step_index * c->tick_den;
      c->cur_op = {SubstepOp::kAbStep, c->order, t, t + c->tick_den, true};
    }
    if (c->free_slots.empty()) return fail("internal error: no free history slot");
    c->cur_slot = c->free_slots.back();
    c->free_slots.pop_back();
    c->dt_last = c->dt_slots[c->cur_slot];
    *time = time_at_tick(c, c->cur_op.tick);
  } else if (is_tableau_stepper(c->stepper)) {
    // RungeKutta::next_time_id (RungeKutta.cpp:37-59): substep k > 0 is at
    // t + dt * substep_times[k-1]
    const ButcherTableau& tab = butcher_tableau(c->stepper);
    const double frac = c->rk_substep == 0 ? 0.0 : tab.substep_times[c->rk_substep - 1];
    c->cur_slot = c->rk_substep;
    c->dt_last = c->dt_slots[c->cur_slot];
    // TimeStepId::next_substep: (1 - frac) t_step + frac t_next_step
    if (c->steps_per_slab == 0)
      *time = c->t0 + ((double)c->step_index + frac) * c->dt;
    else if (c->rk_substep == 0)
      *time = time_at_tick(c, c->step_index);
    else
      *time = (1.0 - frac) * time_at_tick(c, c->step_index) +
              frac * time_at_tick(c, c->step_index + 1);
  } else {
    const long long base = c->step_index * 2;
    const long long off[3] = {0, 2, 1};  // substep times t, t+dt, t+dt/2
    c->cur_slot = 0;
    c->dt_last = c->dt_slots[0];
    if (c->steps_per_slab == 0) {
      *time = c->t0 + ((double)(base + off[c->rk_substep]) / 2.0) * c->dt;
    } else {
      const double frac = 0.5 * (double)off[c->rk_substep];
      const double ts = time_at_tick(c, base), te = time_at_tick(c, base + 2);
      *time = c->rk_substep == 0 ? ts : (1.0 - frac) * ts + frac * te;
    }
  }
  c->in_substep = true;
  return prepare_fused_update(c);
}

int dgrhs_set_exponential_filter(dgrhs_ctx* c, int enable, double alpha, int half_power) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (!enable) {
    if (c->filterF) cudaFree(c->filterF);
    c->filterF = nullptr;
    return 0;
  }
  if (half_power < 1) return fail("half_power must be positive");
  std::vector<double> F;
  exponential_filter_matrix(c->N, alpha, (unsigned)half_power, F);
  if (!c->filterF && dev_alloc(&c->filterF, F.size())) return 1;
  CU(h2d_table(c->filterF, F.data(), F.size() * 8));
  std::memcpy(c->filterF_host, F.data(), F.size() * 8);
  return 0;
}

int dgrhs_apply_exponential_filter(dgrhs_ctx* c) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (!c->filterF) return fail("no exponential filter set");
  if (c->in_substep) return fail("apply_exponential_filter inside a substep");
  return apply_filter(c);
}

int dgrhs_exponential_filter_matrix(int N, double alpha, int half_power, double* matrix) {
  if (N < 2 || half_power < 1) return fail("bad arguments");
  std::vector<double> F;
  exponential_filter_matrix(N, alpha, (unsigned)half_power, F);
  std::memcpy(matrix, F.data(), F.size() * 8);
  return 0;
}

int dgrhs_set_split_volume(dgrhs_ctx* c, int enable) {
  CHECK_CTX(c);
  if (enable < 0 || enable > 2) return fail("volume variant must be 0, 1 or 2");
  c->volume_variant = enable;
  return 0;
}

int dgrhs_set_demand_outgoing_char_speeds(dgrhs_ctx* c, int enable) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (c->system != DGRHS_SYSTEM_GH)
    return fail("DemandOutgoingCharSpeeds is a GeneralizedHarmonic boundary condition");
  if (enable && !c->violations) {
    CU(cudaMalloc(&c->violations, 2 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(c->violations, 0, 2 * sizeof(unsigned long long), c->stream));
  } else if (!enable && c->violations) {
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->violations);
    c->violations = nullptr;
  }
  return 0;
}

int dgrhs_check_outgoing_char_speeds(dgrhs_ctx* c, long long* n_violations,
                                     double* min_speed) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (!c->violations) return fail("DemandOutgoingCharSpeeds is not enabled");
  unsigned long long h[2];
  CU(cudaMemcpyAsync(h, c->violations, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  double mn = 0.0;
  std::memcpy(&mn, &h[1], 8);
  if (n_violations) *n_violations = (long long)h[0];
  if (min_speed) *min_speed = h[0] ? mn : 0.0;
  if (h[0])
    return fail("DemandOutgoingCharSpeeds boundary condition violated at %llu face points, "
                "most ingoing speed: %.17g", h[0], mn);
  return 0;
}

int dgrhs_set_fused_update(dgrhs_ctx* c, int enable) {
  CHECK_CTX(c);
  if (c->in_substep) return fail("cannot change the update mode inside a substep");
  c->fuse_update = enable != 0;
  return 0;
}

int dgrhs_end_substep(dgrhs_ctx* c, int* is_step_done) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (!c->in_substep) return fail("end_substep without begin_substep");
  c->in_substep = false;
  int done = 0;
  if (c->stepper == DGRHS_STEPPER_ADAMS_BASHFORTH) {
    const SubstepOp op = c->cur_op;
    c->history.push_back({op.tick, c->cur_slot});  // RecordTimeStepperData
    if (op.kind == SubstepOp::kAbEvalOnly) {
      ab_clean(c, op.order + 1);  // history order was bumped, UpdateU skipped
    } else {
      if (c->upd_active) {
        std::swap(c->u, c->u_alt);  // UpdateU was fused into the volume kernel
      } else if (ab_update(c, op.order, op.tick, op.tick_end)) {  // UpdateU
        return 1;
      }
      ab_clean(c, op.order);                                       // CleanHistory
      if (op.regular) {
        ++c->step_index;
        done = 1;
      }
    }
  } else if (is_tableau_stepper(c->stepper)) {
    // RungeKutta.cpp:69-122 (compute_substep): u = u_start + dt sum_i coef_i f_i
    const ButcherTableau& tab = butcher_tableau(c->stepper);
    const int nsub = (int)tab.result_coefficients.size();
    const int k = c->rk_substep;
    if (k == 0)
      CU(cudaMemcpyAsync(c->u0, c->u, c->state_len() * 8, cudaMemcpyDeviceToDevice,
                         c->stream));
    const std::vector<double>& row =
        k == nsub - 1 ? tab.result_coefficients : tab.substep_coefficients[k];
    std::vector<double> coef{1.0};
    std::vector<const double*> v{c->u0};
    for (size_t i = 0; i < row.size(); ++i)
      if (row[i] != 0.0) {
        coef.push_back(row[i] * c->dt);
        v.push_back(c->dt_slots[i]);
      }
    if (lincomb(c, c->u, 0.0, coef, v)) return 1;
    if (k == nsub - 1) {
      c->rk_substep = 0;
      ++c->step_index;
      done = 1;
    } else {
      c->rk_substep = k + 1;
    }
  } else {
    const double dt = c->dt;
    const double* F = c->dt_slots[0];
    // Rk3HesthavenSsp.cpp:63-81
    if (c->upd_active) {
      // the fused kernel wrote the substep result to u_alt
      if (c->rk_substep == 0) {
        double* old_u = c->u;  // becomes the saved step-start value, no copy
        c->u = c->u_alt;
        c->u_alt = c->u0;
        c->u0 = old_u;
      } else {
        std::swap(c->u, c->u_alt);
      }
      if (c->rk_substep == 2) {
        ++c->step_index;
        done = 1;
      }
      c->rk_substep = (c->rk_substep + 1) % 3;
    } else if (c->rk_substep == 0) {
      CU(cudaMemcpyAsync(c->u0, c->u, c->state_len() * 8, cudaMemcpyDeviceToDevice,
                         c->stream));
      if (lincomb(c, c->u, 1.0, {dt}, {F})) return 1;
      c->rk_substep = 1;
    } else if (c->rk_substep == 1) {
      if (lincomb(c, c->u, 0.25, {0.75, 0.25 * dt}, {c->u0, F})) return 1;
      c->rk_substep = 2;
    } else {
      if (lincomb(c, c->u, 2.0 / 3.0, {1.0 / 3.0, (2.0 / 3.0) * dt}, {c->u0, F})) return 1;
      c->rk_substep = 0;
      ++c->step_index;
      done = 1;
    }
  }
  // dg::Actions::Filter runs after UpdateU in step_actions; a self-start
  // substep whose update is skipped leaves u untouched (and is reset anyway)
  const bool updated = !(c->stepper == DGRHS_STEPPER_ADAMS_BASHFORTH &&
                         c->cur_op.kind == SubstepOp::kAbEvalOnly);
  if (updated && apply_filter(c)) return 1;
  c->upd_active = false;
  if (is_step_done) *is_step_done = done;
  return 0;
}

// ---------------------------------------------------------------------------
// Halo exchange inside the library: NCCL send/recv of the cut mortar faces
// (reference: send_data_for_fluxes / receive_boundary_data_global_time_stepping,
// ComputeTimeDerivative.hpp:652-774, ApplyBoundaryCorrections.hpp:205-380).
// libnccl is resolved at run time (dlopen) so that single-GPU callers need no
// NCCL and a process that already loaded a libnccl.so.2 (e.g. PyTorch's) shares it.
// ---------------------------------------------------------------------------
}  // extern "C"

#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  if (api.handle) return &api;
  void* h = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    fail("cannot load libnccl.so.2: %s", dlerror());
    return nullptr;
  }
#define DG_NCCL_SYM(field, sym)                                   \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym)); \
  if (!api.field) {                                               \
    fail("libnccl lacks %s", sym);                                \
    return nullptr;                                               \
  }
  DG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  DG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  DG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  DG_NCCL_SYM(Send, "ncclSend")
  DG_NCCL_SYM(Recv, "ncclRecv")
  DG_NCCL_SYM(GroupStart, "ncclGroupStart")
  DG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  DG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef DG_NCCL_SYM
  api.handle = h;
  return &api;
}

#define NC(call)                                                                       \
  do {                                                                                 \
    ncclResult_t r__ = (call);                                                         \
    if (r__ != ncclSuccess)                                                            \
      return fail("%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r__), __FILE__, \
                  __LINE__);                                                           \
  } while (0)

// queue the exchange of the packed faces on the communication stream
int start_halo_exchange(dgrhs_ctx* c) {
  NcclApi* nc = nccl_api();
  if (!nc) return 1;
  CU(cudaEventRecord(c->ev_packed, c->stream));
  CU(cudaStreamWaitEvent(c->comm_stream, c->ev_packed, 0));
  if (c->phase_timing) CU(cudaEventRecord(c->phase_ev[4], c->comm_stream));
  const size_t per_face = (size_t)c->HC * c->f;
  ncclComm_t comm = static_cast<ncclComm_t>(c->nccl_comm);
  size_t so = 0, ro = 0;
  NC(nc->GroupStart());
  for (int peer = 0; peer < c->comm_world; ++peer) {
    const size_t ns = (size_t)c->send_counts[peer], nr = (size_t)c->recv_counts[peer];
    if (nr)
      NC(nc->Recv(c->halo_recv + ro * per_face, nr * per_face, ncclDouble, peer, comm,
                  c->comm_stream));
    if (ns)
      NC(nc->Send(c->halo_send + so * per_face, ns * per_face, ncclDouble, peer, comm,
                  c->comm_stream));
    so += ns;
    ro += nr;
  }
  NC(nc->GroupEnd());
  CU(cudaEventRecord(c->ev_halo, c->comm_stream));
  if (c->phase_timing) CU(cudaEventRecord(c->phase_ev[5], c->comm_stream));
  return 0;
}

// One RHS evaluation of a rank that exchanges faces.  Streams:
//   main: pack | faces of the interfaces that touch an interior element | volume of the
//         interior elements | (join)
//   comm: (packed) NCCL send/recv | faces of the remaining interfaces (they need the halo and
//         write corrections of boundary elements only) | (faces of main done) volume of the
//         boundary elements
// The boundary pass runs NEXT TO the interior volume kernel (higher stream priority), not
// behind it: no kernel tail of the interior pass is waited for and no launch gap is exposed.
int rhs_with_exchange(dgrhs_ctx* c, double t) {
  const dg::UpdateArgs upd = c->upd_active ? c->pending_upd : dg::UpdateArgs{};
  const DgNOps* ops = dgrhs_nops(c->N);
  cudaStream_t main_stream = c->stream;
  const bool pt = c->phase_timing;
  if (pt) CU(cudaEventRecord(c->phase_ev[0], main_stream));
  // pack + exchange on the communication stream, behind everything queued so far
  CU(cudaEventRecord(c->ev_packed, main_stream));
  CU(cudaStreamWaitEvent(c->comm_stream, c->ev_packed, 0));
  c->stream = c->comm_stream;
  int prc = ops->pack(c);
  if (!prc && pt && cudaEventRecord(c->phase_ev[1], c->comm_stream) != cudaSuccess) prc = 1;
  if (!prc) prc = start_halo_exchange(c);
  c->stream = main_stream;
  if (prc) return 1;
  ++c->rhs_evals;
  const int ni = c->n_interior;
  c->pdl_volume = false;
  if (ops->gauge(c, t)) return 1;
  if (ni > 0 && ops->faces(c, 0, ni)) return 1;
  if (c->bjorhus_join_pending) {  // (every element is interior: whole-batch pass)
    c->bjorhus_join_pending = false;
    CU(cudaStreamWaitEvent(main_stream, c->aux_join, 0));
  }
  CU(cudaEventRecord(c->ev_faces1, main_stream));
  if (pt) CU(cudaEventRecord(c->phase_ev[2], main_stream));
  c->pdl_volume = false;  // an event sits between the faces and the volume kernel
  if (ni > 0 && ops->volume(c, c->dt_last, 0, ni, true, &upd)) return 1;
  if (c->mesh_v && ops->mesh_velocity_terms(c, c->dt_last, 0, ni)) return 1;
  if (pt) CU(cudaEventRecord(c->phase_ev[3], main_stream));
  {
    c->stream = c->comm_stream;  // the launchers queue on c->stream
    int rc = ops->faces(c, ni, c->nelem);
    if (pt) cudaEventRecord(c->phase_ev[6], c->comm_stream);
    c->pdl_volume = false;
    if (!rc && cudaStreamWaitEvent(c->comm_stream, c->ev_faces1, 0) != cudaSuccess) rc = 1;
    if (!rc) rc = ops->volume(c, c->dt_last, ni, c->nelem, true, &upd);
    if (!rc && c->mesh_v) rc = ops->mesh_velocity_terms(c, c->dt_last, ni, c->nelem);
    c->stream = main_stream;
    if (rc) return 1;
  }
  if (pt) CU(cudaEventRecord(c->phase_ev[7], c->comm_stream));
  CU(cudaEventRecord(c->ev_faces2, c->comm_stream));
  CU(cudaStreamWaitEvent(main_stream, c->ev_faces2, 0));
  return 0;
}

}  // namespace

extern "C" {

void dgrhs_internal_comm_destroy(void* comm) {
  if (NcclApi* nc = nccl_api()) nc->CommDestroy(static_cast<ncclComm_t>(comm));
}

int dgrhs_comm_unique_id(void* unique_id_128_bytes) {
  NcclApi* nc = nccl_api();
  if (!nc) return 1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NC(nc->GetUniqueId(static_cast<ncclUniqueId*>(unique_id_128_bytes)));
  return 0;
}

int dgrhs_comm_init(dgrhs_ctx* c, const void* unique_id_128_bytes, int rank, int world) {
  CHECK_CTX(c);
  if (world < 1 || rank < 0 || rank >= world) return fail("bad rank %d of %d", rank, world);
  if (c->nccl_comm) return fail("communicator already initialised");
  NcclApi* nc = nccl_api();
  if (!nc) return 1;
  CU(cudaSetDevice(c->device));
  ncclUniqueId id;
  std::memcpy(&id, unique_id_128_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  NC(nc->CommInitRank(&comm, world, id, rank));
  c->nccl_comm = comm;
  c->comm_rank = rank;
  c->comm_world = world;
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CU(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
  CU(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_faces2, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_faces1, cudaEventDisableTiming));
  c->send_counts.assign(world, 0);
  c->recv_counts.assign(world, 0);
  return 0;
}

int dgrhs_set_halo_peers(dgrhs_ctx* c, const int32_t* send_counts, const int32_t* recv_counts) {
  CHECK_CTX(c);
  if (!c->nccl_comm) return fail("dgrhs_comm_init has not been called");
  long long ns = 0, nr = 0;
  for (int p = 0; p < c->comm_world; ++p) {
    if (send_counts[p] < 0 || recv_counts[p] < 0) return fail("negative face count");
    if (p == c->comm_rank && (send_counts[p] || recv_counts[p]))
      return fail("a rank does not exchange faces with itself");
    ns += send_counts[p];
    nr += recv_counts[p];
  }
  if (ns != c->n_send) return fail("send counts sum to %lld, the halo map has %d faces", ns, c->n_send);
  if (nr > c->nghost) return fail("recv counts sum to %lld, the context has %d ghost slots", nr, c->nghost);
  c->send_counts.assign(send_counts, send_counts + c->comm_world);
  c->recv_counts.assign(recv_counts, recv_counts + c->comm_world);
  return 0;
}

int dgrhs_set_phase_timing(dgrhs_ctx* c, int enable) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (enable && !c->phase_ev[0])
    for (auto& e : c->phase_ev) CU(cudaEventCreate(&e));
  c->phase_timing = enable != 0;
  return 0;
}

int dgrhs_get_phase_times(dgrhs_ctx* c, double* ms) {
  CHECK_CTX(c);
  if (!c->phase_timing || !c->nccl_comm) return fail("phase timing is not enabled");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaStreamSynchronize(c->comm_stream));
  for (int i = 1; i < 8; ++i) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, c->phase_ev[0], c->phase_ev[i]));
    ms[i - 1] = t;
  }
  return 0;
}

int dgrhs_exchange_halo(dgrhs_ctx* c) {
  CHECK_CTX(c);
  if (!c->nccl_comm) return fail("dgrhs_comm_init has not been called");
  CU(cudaSetDevice(c->device));
  if (dgrhs_nops(c->N)->pack(c)) return 1;
  if (start_halo_exchange(c)) return 1;
  CU(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
  return 0;
}

int dgrhs_take_steps(dgrhs_ctx* c, int n_steps) {
  CHECK_CTX(c);
  const bool exchange = c->n_send > 0 || (c->nccl_comm && c->comm_world > 1);
  if (exchange) {
    if (!c->nccl_comm)
      return fail("context exchanges faces with other ranks: call dgrhs_comm_init + "
                  "dgrhs_set_halo_peers, or drive the substeps from the caller");
    if (c->n_interior < 0) return fail("dgrhs_set_interior_count has not been called");
  }
  for (int s = 0; s < n_steps;) {
    double t;
    int done = 0;
    if (dgrhs_begin_substep(c, &t)) return 1;
    if (exchange) {
      if (rhs_with_exchange(c, t)) return 1;
    } else {
      ++c->rhs_evals;
      if (rhs_range(c, t, c->dt_last, 0, c->nelem, false, true,
                    c->upd_active ? c->pending_upd : dg::UpdateArgs{}))
        return 1;
    }
    if (dgrhs_end_substep(c, &done)) return 1;
    if (done) ++s;
  }
  return 0;
}

double dgrhs_time(dgrhs_ctx* c) { return c ? time_at_tick(c, c->step_index * c->tick_den) : 0.0; }
int64_t dgrhs_rhs_evaluations(dgrhs_ctx* c) { return c ? c->rhs_evals : 0; }

int dgrhs_time_kernels(dgrhs_ctx* c, int reps, int update_terms, double* ms) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  if (reps < 1 || update_terms < 1 || update_terms > 6) return fail("bad arguments");
  if (ensure_slots(c, update_terms + 1)) return 1;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double* scratch = c->dt_slots[update_terms];
  if (!c->u_alt && dev_alloc(&c->u_alt, c->state_len())) return 1;
  dg::UpdateArgs fused{};
  fused.u_new = c->u_alt;
  fused.a = 1.0;
  fused.c_new = 0.0;
  fused.nterms = std::min(update_terms - 1, 3);
  for (int j = 0; j < fused.nterms; ++j) {
    fused.c[j] = 0.0;
    fused.v[j] = c->dt_slots[j];
  }
  for (int which = 0; which < 4; ++which) {
    float total = 0.f;
    for (int r = -1; r < reps; ++r) {  // r = -1: warm-up
      CU(cudaEventRecord(e0, c->stream));
      int rc = 0;
      const DgNOps* ops = dgrhs_nops(c->N);
      if (which == 0) {
        c->aux_faces_eval = -1;  // time the Bjorhus/mortar kernels too
        rc = ops->faces(c, 0, c->nelem);
        if (c->bjorhus_join_pending) {  // no volume launch follows here: join now
          c->bjorhus_join_pending = false;
          CU(cudaStreamWaitEvent(c->stream, c->aux_join, 0));
        }
      }
      if (which == 1) rc = ops->volume(c, c->dt_last, 0, c->nelem, true, nullptr);
      if (which == 3) rc = ops->volume(c, scratch, 0, c->nelem, true, &fused);
      if (which == 2) {
        std::vector<double> coef(update_terms, 0.0);
        std::vector<const double*> v;
        for (int j = 0; j < update_terms; ++j) v.push_back(c->dt_slots[j]);
        rc = lincomb(c, scratch, 1.0, coef, v);
      }
      if (rc) return 1;
      CU(cudaEventRecord(e1, c->stream));
      CU(cudaEventSynchronize(e1));
      float t;
      CU(cudaEventElapsedTime(&t, e0, e1));
      if (r >= 0) total += t;
    }
    ms[which] = total / reps;
  }
  // the filter pass (apply_matrices with three N x N matrices per component), on the
  // scratch state so that u stays untouched
  ms[4] = 0.0;
  if (c->filterF) {
    double* keep = c->u;
    c->u = c->u_alt;
    float total = 0.f;
    int rc = 0;
    for (int r = -1; r < reps && !rc; ++r) {
      cudaEventRecord(e0, c->stream);
      rc = apply_filter(c);
      cudaEventRecord(e1, c->stream);
      cudaEventSynchronize(e1);
      float t;
      cudaEventElapsedTime(&t, e0, e1);
      if (r >= 0) total += t;
    }
    c->u = keep;
    if (rc) return 1;
    ms[4] = total / reps;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

int dgrhs_gh_constraint_norms(dgrhs_ctx* c, double* norms) {
  CHECK_CTX(c);
  if (c->system != DGRHS_SYSTEM_GH) return fail("constraint norms are defined for GH only");
  if (c->gauge == DGRHS_GAUGE_DAMPED_HARMONIC)
    return fail("gauge-constraint norm with the DampedHarmonic gauge is not implemented");
  CU(cudaSetDevice(c->device));
  double* sums = nullptr;
  if (dev_alloc(&sums, 3)) return 1;
  if (dgrhs_nops(c->N)->constraints(c, sums)) return 1;
  double h[3];
  CU(cudaMemcpyAsync(h, sums, 24, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(sums);
  const double npts = (double)c->nelem * c->n;
  for (int i = 0; i < 3; ++i) norms[i] = std::sqrt(h[i] / npts);
  return 0;
}

int dgrhs_synchronize(dgrhs_ctx* c) {
  CHECK_CTX(c);
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
void* dgrhs_stream(dgrhs_ctx* c) { return c ? (void*)c->stream : nullptr; }
void* dgrhs_state_device_ptr(dgrhs_ctx* c) { return c ? c->u : nullptr; }
int dgrhs_padded_points(dgrhs_ctx* c) { return c ? c->npad : 0; }

// ---- single-operator entry points ----------------------------------------

int dgrhs_differentiation_matrix(int N, double* matrix) {
  if (N < 2) return fail("need at least two LGL points");
  std::vector<double> D;
  diff_matrix(N, D);
  std::memcpy(matrix, D.data(), D.size() * 8);
  return 0;
}

int dgrhs_collocation_points_and_weights(int N, double* points, double* weights) {
  if (N < 2) return fail("need at least two LGL points");
  std::vector<double> x, w;
  lgl(N, x, w);
  std::memcpy(points, x.data(), N * 8);
  std::memcpy(weights, w.data(), N * 8);
  return 0;
}

int dgrhs_adams_bashforth_coefficients(int order, const double* times, double step_start,
                                       double step_end, double* coefficients) {
  if (order < 1 || order > 8) return fail("order must be in [1, 8]");
  const double step = step_end - step_start;
  bool constant = true;
  std::vector<double> control{0.0};
  for (int i = 1; i < order; ++i) {
    const double this_step = times[i] - times[i - 1];
    control.push_back(control.back() + this_step);
    // slab_rounding_error: a few ulp of the larger of |t| and the step
    if (std::abs(this_step - step) >
        4.0 * 2.220446049250313e-16 * std::max(std::abs(times[i]), std::abs(step)))
      constant = false;
  }
  std::vector<double> r;
  if (constant && step_start == times[order - 1]) {
    for (int i = 0; i < order; ++i) r.push_back(kAbConst[order][i] * step);
  } else {
    r = variable_coefficients(control, control.back() + (step_start - times[order - 1]),
                              control.back() + (step_end - times[order - 1]));
  }
  std::memcpy(coefficients, r.data(), order * 8);
  return 0;
}

// TimeStepper::order / number_of_substeps / number_of_past_steps / stable_step.
// The stable step (largest dt for which y' = -2 y / ... decays, normalised so that
// forward Euler gives 1) is computed from first principles: Adams-Bashforth from the
// root zeta = -1 of the characteristic polynomial, Runge-Kutta from the first
// crossing |R(-2 x)| = 1 of the stability polynomial R(z) = 1 + sum_k z^k b.A^(k-1).1
// of the tableau -- the CPU tests compare with the constants of the reference
// (AdamsBashforth.cpp:72-90, Rk3HesthavenSsp.cpp:23-26, Rk3Owren.cpp:15,
// Rk3Kennedy.cpp:10, ClassicalRungeKutta4.cpp:22, DormandPrince5.cpp:19).
int dgrhs_butcher_row(int stepper, int substep, double* coefficients) {
  if (!is_tableau_stepper(stepper)) return fail("stepper %d has no Butcher tableau", stepper);
  const ButcherTableau& tab = butcher_tableau(stepper);
  const int nsub = (int)tab.result_coefficients.size();
  if (substep < 0 || substep >= nsub) return fail("substep out of range");
  const std::vector<double>& row =
      substep == nsub - 1 ? tab.result_coefficients : tab.substep_coefficients[substep];
  for (int i = 0; i <= substep; ++i) coefficients[i] = i < (int)row.size() ? row[i] : 0.0;
  return 0;
}

int dgrhs_stepper_substep_fractions(int stepper, double* fractions) {
  if (stepper == DGRHS_STEPPER_ADAMS_BASHFORTH) return 0;
  if (stepper == DGRHS_STEPPER_RK3_HESTHAVEN) {
    fractions[0] = 1.0;
    fractions[1] = 0.5;
    return 0;
  }
  if (!is_tableau_stepper(stepper)) return fail("unknown time stepper %d", stepper);
  const ButcherTableau& tab = butcher_tableau(stepper);
  const size_t nsub = tab.result_coefficients.size();
  for (size_t k = 0; k + 1 < nsub; ++k) fractions[k] = tab.substep_times[k];
  return 0;
}

int dgrhs_stepper_properties(int stepper, int order, int* order_out, int* number_of_substeps,
                             int* number_of_past_steps, double* stable_step) {
  int ord = 0, substeps = 0, past = 0;
  double stable = 0.0;
  if (stepper == DGRHS_STEPPER_ADAMS_BASHFORTH) {
    if (order < 1 || order > 8) return fail("AdamsBashforth order must be in [1, 8]");
    ord = order;
    substeps = 1;
    past = order - 1;
    double alternating = 0.0;  // sum_j beta_j (-1)^j, j = 0 the newest derivative
    for (int j = 0; j < order; ++j)
      alternating += kAbConst[order][order - 1 - j] * ((j & 1) ? -1.0 : 1.0);
    stable = 1.0 / alternating;
  } else {
    std::vector<double> poly{1.0};  // R(z) coefficients, lowest degree first
    if (stepper == DGRHS_STEPPER_RK3_HESTHAVEN) {
      ord = 3;
      substeps = 3;
      poly = {1.0, 1.0, 0.5, 1.0 / 6.0};  // any 3-stage method of order 3
    } else if (is_tableau_stepper(stepper)) {
      const ButcherTableau& tab = butcher_tableau(stepper);
      const int s = (int)tab.result_coefficients.size();
      substeps = s;
      ord = stepper == DGRHS_STEPPER_RK4 ? 4 : stepper == DGRHS_STEPPER_DORMAND_PRINCE5 ? 5 : 3;
      std::vector<double> v(s, 1.0);  // A^(k-1) 1
      for (int k = 1; k <= s; ++k) {
        double bk = 0.0;
        for (int i = 0; i < s; ++i) bk += tab.result_coefficients[i] * v[i];
        poly.push_back(bk);
        std::vector<double> w(s, 0.0);
        for (int i = 1; i < s; ++i)
          for (int j = 0; j < i; ++j) w[i] += tab.substep_coefficients[i - 1][j] * v[j];
        v = w;
      }
    } else {
      return fail("unknown time stepper %d", stepper);
    }
    auto growth = [&](double x) {
      double r = 0.0;
      for (size_t k = poly.size(); k-- > 0;) r = r * (-2.0 * x) + poly[k];
      return std::abs(r) - 1.0;
    };
    double lo = 1e-6, hi = lo;
    while (growth(hi) < 0.0 && hi < 100.0) {
      lo = hi;
      hi += 1e-3;
    }
    for (int it = 0; it < 200; ++it) {
      const double mid = 0.5 * (lo + hi);
      (growth(mid) < 0.0 ? lo : hi) = mid;
    }
    stable = 0.5 * (lo + hi);
  }
  if (order_out) *order_out = ord;
  if (number_of_substeps) *number_of_substeps = substeps;
  if (number_of_past_steps) *number_of_past_steps = past;
  if (stable_step) *stable_step = stable;
  return 0;
}

int dgrhs_partial_derivatives(int N, int C, const double* u, const double* invjac,
                              double* du) {
  if (N < 2 || N > 12) return fail("n_points_1d must be in [2, 12]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device available: this library has no CPU fallback");
  const int n = N * N * N, npad = dg::padded_points(N);
  double *du_d = nullptr, *u_d = nullptr, *j_d = nullptr, *D_d = nullptr;
  if (dev_alloc(&u_d, (size_t)C * npad) || dev_alloc(&j_d, (size_t)9 * npad) ||
      dev_alloc(&du_d, (size_t)3 * C * npad) || dev_alloc(&D_d, (size_t)N * N))
    return 1;
  std::vector<double> D;
  diff_matrix(N, D);
  CU(h2d_table(D_d, D.data(), D.size() * 8));
  CU(cudaMemcpy2D(u_d, (size_t)npad * 8, u, (size_t)n * 8, (size_t)n * 8, C,
                  cudaMemcpyHostToDevice));
  CU(cudaMemcpy2D(j_d, (size_t)npad * 8, invjac, (size_t)n * 8, (size_t)n * 8, 9,
                  cudaMemcpyHostToDevice));
  dg::DerivArgs a{u_d, j_d, du_d, D_d, C, 3 * C, 0, 3};
  if (dgrhs_nops(N)->partial_derivatives(&a, C, nullptr)) return 1;
  CU(cudaMemcpy2D(du, (size_t)n * 8, du_d, (size_t)npad * 8, (size_t)n * 8, 3 * C,
                  cudaMemcpyDeviceToHost));
  cudaFree(u_d);
  cudaFree(j_d);
  cudaFree(du_d);
  cudaFree(D_d);
  return 0;
}

}  // extern "C"
