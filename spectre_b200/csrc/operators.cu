// Single-operator entry points of include/dgrhs.h: GPU forwards with the
// argument meaning of the reference's per-element operator surface
// (TimeDerivative::apply, UpwindPenalty::dg_package_data / dg_boundary_terms,
// lift_flux).  Host pointers in, host pointers out; the data make one round
// trip over PCIe per call, so these are for drop-in use at operator
// granularity and for parity tests that read like the reference's own unit
// tests -- the batched context API is the fast path.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/dgrhs.h"
#include "pointwise.cuh"
#include "bjorhus.cuh"

extern "C" void dgrhs_internal_set_error(const char* msg);
extern "C" void dgrhs_internal_count_launch(void);

namespace {

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  dgrhs_internal_set_error(buf);
  return 1;
}

#define CU(call)                                                             \
  do {                                                                       \
    cudaError_t err__ = (call);                                              \
    if (err__ != cudaSuccess)                                                \
      return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                       \
  } while (0)

int need_gpu() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device available: this library has no CPU fallback");
  return 0;
}

// RAII device staging of host arrays
struct Staged {
  std::vector<void*> ptrs;
  ~Staged() {
    for (void* p : ptrs) cudaFree(p);
  }
  double* in(const double* h, size_t count) {
    double* d = nullptr;
    if (cudaMalloc((void**)&d, std::max<size_t>(count, 1) * 8) != cudaSuccess) return nullptr;
    ptrs.push_back(d);
    if (h && cudaMemcpy(d, h, count * 8, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    return d;
  }
};

// gh::TimeDerivative<3>::apply on given derivatives (TimeDerivative.cpp:31-407)
template <bool kHarmonic>
__global__ void gh_time_derivative_kernel(int n, const double* u, const double* du,
                                          const double* g0, const double* g1,
                                          const double* g2, const double* H,
                                          const double* dH, double* dt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double g[10], pi[10], phi[3][10], J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, Q[10];
#pragma unroll
  for (int s = 0; s < 10; ++s) {
    g[s] = u[(size_t)s * n + p];
    pi[s] = u[(size_t)(10 + s) * n + p];
#pragma unroll
    for (int m = 0; m < 3; ++m) phi[m][s] = u[(size_t)(20 + m + 3 * s) * n + p];
  }
  dg::GaugeH gh;
  if (!kHarmonic) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      gh.H[a] = H[(size_t)a * n + p];
#pragma unroll
      for (int b = 0; b < 4; ++b) gh.dH[a][b] = dH[(size_t)(a + 4 * b) * n + p];
    }
  }
  dg::GhContext ctx;
  dg::GaugeInput gin;
  gin.fields = &gh;
  dg::gh_prologue<kHarmonic ? 0 : 1>(g, pi, phi, J, g0[p], g1[p], g2[p], gin, ctx, Q);
#pragma unroll 1
  for (int s = 0; s < 10; ++s) {
    double ph[3], dgv[3], dpi[3], dph[3][3], og, op, oph[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) ph[m] = u[(size_t)(20 + m + 3 * s) * n + p];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dgv[i] = du[(size_t)(3 * s + i) * n + p];
      dpi[i] = du[(size_t)(3 * (10 + s) + i) * n + p];
#pragma unroll
      for (int m = 0; m < 3; ++m) dph[m][i] = du[(size_t)(3 * (20 + m + 3 * s) + i) * n + p];
    }
    // Q is indexed dynamically: select from the register array
    double Qs = Q[0];
#pragma unroll
    for (int t = 1; t < 10; ++t) Qs = (s == t) ? Q[t] : Qs;
    dg::gh_pair_rhs(ctx, Qs, u[(size_t)s * n + p], u[(size_t)(10 + s) * n + p], ph, dgv, dpi,
                    dph, og, op, oph);
    dt[(size_t)s * n + p] = og;
    dt[(size_t)(10 + s) * n + p] = op;
#pragma unroll
    for (int m = 0; m < 3; ++m) dt[(size_t)(20 + m + 3 * s) * n + p] = oph[m];
  }
}

__global__ void sw_time_derivative_kernel(int n, const double* u, const double* du,
                                          const double* gamma2, double* dt) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double up[5], d[5][3], out[5];
  const double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    up[c] = u[(size_t)c * n + p];
#pragma unroll
    for (int i = 0; i < 3; ++i) d[c][i] = du[(size_t)(3 * c + i) * n + p];
  }
  dg::sw_point_rhs(up, d, J, gamma2[p], out);
#pragma unroll
  for (int c = 0; c < 5; ++c) dt[(size_t)c * n + p] = out[c];
}

// gh UpwindPenalty::dg_package_data with lapse, shift and both normals given
// (UpwindPenalty.cpp:36-158); packaged order of dg_package_field_tags
__global__ void gh_package_kernel(int f, const double* u, const double* g1,
                                  const double* g2, const double* lapse,
                                  const double* shift, const double* n_lo,
                                  const double* n_up, const double* ndotv, double* pk,
                                  double* max_speed) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= f) return;
  dg::GhFaceSide s;
  double sdn = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    s.n_lo[i] = n_lo[(size_t)i * f + p];
    s.n_up[i] = n_up[(size_t)i * f + p];
    sdn += shift[(size_t)i * f + p] * s.n_lo[i];
  }
  sdn = -sdn;
  s.speed[1] = sdn;
  s.speed[0] = (1.0 + g1[p]) * sdn;
  s.speed[2] = lapse[p] + sdn;
  s.speed[3] = -lapse[p] + sdn;
  if (ndotv) {  // moving mesh (UpwindPenalty.cpp:85-91)
    s.speed[0] -= ndotv[p] * (1.0 + g1[p]);
    s.speed[1] -= ndotv[p];
    s.speed[2] -= ndotv[p];
    s.speed[3] -= ndotv[p];
  }
  s.gamma2 = g2[p];
  s.mag = 1.0;
#pragma unroll 1
  for (int a = 0; a < 10; ++a) {
    double ph[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) ph[m] = u[(size_t)(20 + m + 3 * a) * f + p];
    dg::GhPairPackaged k;
    dg::gh_pair_package(s, u[(size_t)a * f + p], u[(size_t)(10 + a) * f + p], ph, k);
    pk[(size_t)a * f + p] = k.v_g;
    pk[(size_t)(40 + a) * f + p] = k.v_plus;
    pk[(size_t)(50 + a) * f + p] = k.v_minus;
    pk[(size_t)(120 + a) * f + p] = k.g2_v_g;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      pk[(size_t)(10 + m + 3 * a) * f + p] = k.v_zero[m];
      pk[(size_t)(60 + m + 3 * a) * f + p] = k.v_plus * s.n_lo[m];
      pk[(size_t)(90 + m + 3 * a) * f + p] = k.v_minus * s.n_lo[m];
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) pk[(size_t)(130 + c) * f + p] = s.speed[c];
  max_speed[p] = fmax(fmax(s.speed[0], s.speed[1]), fmax(s.speed[2], s.speed[3]));
}

// gh UpwindPenalty::dg_boundary_terms (UpwindPenalty.cpp:161-275)
__global__ void gh_boundary_terms_kernel(int f, const double* in, const double* ex,
                                         double* corr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= f) return;
  double wi[4], we[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    wi[c] = dg::step_function(-in[(size_t)(130 + c) * f + p]);
    we[c] = -dg::step_function(ex[(size_t)(130 + c) * f + p]);
  }
#pragma unroll 1
  for (int a = 0; a < 10; ++a) {
    auto I = [&](int c) { return in[(size_t)c * f + p]; };
    auto E = [&](int c) { return ex[(size_t)c * f + p]; };
    corr[(size_t)a * f + p] = we[0] * E(a) - wi[0] * I(a);
    corr[(size_t)(10 + a) * f + p] =
        0.5 * (we[2] * E(40 + a) + we[3] * E(50 + a)) + we[0] * E(120 + a) -
        0.5 * (wi[2] * I(40 + a) + wi[3] * I(50 + a)) - wi[0] * I(120 + a);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int k = d + 3 * a;
      corr[(size_t)(20 + k) * f + p] =
          -0.5 * (we[3] * E(90 + k) - we[2] * E(60 + k)) + we[1] * E(10 + k) -
          0.5 * (wi[2] * I(60 + k) - wi[3] * I(90 + k)) - wi[1] * I(10 + k);
    }
  }
}

// ScalarWave UpwindPenalty (UpwindPenalty.cpp:36-205), packaged order of
// dg_package_field_tags: v_psi, v_zero(3), v_plus, v_minus, n v_plus(3),
// n v_minus(3), gamma2 v_psi, speeds(3)
__global__ void sw_package_kernel(int f, const double* u, const double* gamma2,
                                  const double* n, const double* ndotv, double* pk,
                                  double* max_speed) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= f) return;
  const double psi = u[p], pi = u[(size_t)f + p];
  double phi[3], nn[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    phi[i] = u[(size_t)(2 + i) * f + p];
    nn[i] = n[(size_t)i * f + p];
  }
  const double nv = ndotv ? ndotv[p] : 0.0;  // moving mesh (UpwindPenalty.cpp:55-67)
  const double cs[3] = {0.0 - nv, 1.0 - nv, -1.0 - nv};
  const double g2psi = gamma2[p] * psi;
  double ndphi = nn[0] * phi[0];
  ndphi += nn[1] * phi[1];
  ndphi += nn[2] * phi[2];
  const double vp = cs[1] * (pi + ndphi - g2psi), vm = cs[2] * (pi - ndphi - g2psi);
  pk[p] = cs[0] * psi;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    pk[(size_t)(1 + i) * f + p] = cs[0] * (phi[i] - nn[i] * ndphi);
    pk[(size_t)(6 + i) * f + p] = vp * nn[i];
    pk[(size_t)(9 + i) * f + p] = vm * nn[i];
    pk[(size_t)(13 + i) * f + p] = cs[i];
  }
  pk[(size_t)4 * f + p] = vp;
  pk[(size_t)5 * f + p] = vm;
  pk[(size_t)12 * f + p] = g2psi * cs[0];
  max_speed[p] = fmax(cs[0], fmax(cs[1], cs[2]));
}

__global__ void sw_boundary_terms_kernel(int f, const double* in, const double* ex,
                                         double* corr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= f) return;
  auto I = [&](int c) { return in[(size_t)c * f + p]; };
  auto E = [&](int c) { return ex[(size_t)c * f + p]; };
  const double w0i = dg::step_function(-I(13)), w0e = -dg::step_function(E(13));
  const double wpi = dg::step_function(-I(14)), wpe = -dg::step_function(E(14));
  const double wmi = dg::step_function(-I(15)), wme = -dg::step_function(E(15));
  corr[p] = w0e * E(0) - w0i * I(0);
  corr[(size_t)f + p] = 0.5 * (wpe * E(4) + wme * E(5)) + w0e * E(12) -
                        0.5 * (wpi * I(4) + wmi * I(5)) - w0i * I(12);
#pragma unroll
  for (int d = 0; d < 3; ++d)
    corr[(size_t)(2 + d) * f + p] = 0.5 * (wpe * E(6 + d) - wme * E(9 + d)) + w0e * E(1 + d) -
                                    0.5 * (wpi * I(6 + d) - wmi * I(9 + d)) - w0i * I(1 + d);
}

// dg::lift_flux (LiftFlux.hpp:57-61)
__global__ void lift_flux_kernel(int f, int ncomp, double* corr, int extent,
                                 const double* mag) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= f) return;
  const double s = -0.5 * (double)(extent * (extent - 1)) * mag[p];
  for (int c = 0; c < ncomp; ++c) corr[(size_t)c * f + p] *= s;
}

int grid(int n) { return (n + 127) / 128; }

}  // namespace

// gh::BoundaryConditions::ConstraintPreservingBjorhus<3>::dg_time_derivative
// (Bjorhus.cpp:104-391) on n face points, every tensor in the reference's storage
// order (first index fastest, symmetric pairs 00 01 02 03 11 12 13 22 23 33)
struct BjorhusOpArgs {
  const double *n_lo, *g, *pi, *phi, *x, *gamma1, *gamma2, *lapse, *shift, *ipsi, *t_up, *c3, *H,
      *dH, *dt_g, *dt_pi, *dt_phi, *d_pi, *d_phi;
  double *out_g, *out_pi, *out_phi;
};

__global__ void gh_bjorhus_op_kernel(int n, int physical, BjorhusOpArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  dg::BjorhusInput in;
  in.physical = physical != 0;
  in.gamma1 = a.gamma1[p];
  in.gamma2 = a.gamma2[p];
  in.lapse = a.lapse[p];
  for (int i = 0; i < 3; ++i) {
    in.n_lo[i] = a.n_lo[(size_t)i * n + p];
    in.x[i] = a.x[(size_t)i * n + p];
    in.shift[i] = a.shift[(size_t)i * n + p];
  }
  for (int aa = 0; aa < 4; ++aa) {
    in.t_up[aa] = a.t_up[(size_t)aa * n + p];
    in.H[aa] = a.H[(size_t)aa * n + p];
    for (int bb = 0; bb < 4; ++bb) {
      const size_t s = dg::sym4(aa, bb);
      in.dH[aa][bb] = a.dH[(size_t)(aa + 4 * bb) * n + p];
      in.g[aa][bb] = a.g[s * n + p];
      in.pi[aa][bb] = a.pi[s * n + p];
      in.ipsi[aa][bb] = a.ipsi[s * n + p];
      in.dt_g[aa][bb] = a.dt_g[s * n + p];
      in.dt_pi[aa][bb] = a.dt_pi[s * n + p];
      for (int i = 0; i < 3; ++i) {
        in.phi[i][aa][bb] = a.phi[(i + 3 * s) * n + p];
        in.c3[i][aa][bb] = a.c3[(i + 3 * s) * n + p];
        in.dt_phi[i][aa][bb] = a.dt_phi[(i + 3 * s) * n + p];
        in.d_pi[i][aa][bb] = a.d_pi[(i + 3 * s) * n + p];
        for (int j = 0; j < 3; ++j) in.d_phi[i][j][aa][bb] = a.d_phi[(i + 3 * (j + 3 * s)) * n + p];
      }
    }
  }
  dg::BjorhusOutput out;
  dg::bjorhus_constraint_preserving(in, out);
  for (int aa = 0; aa < 4; ++aa)
    for (int bb = aa; bb < 4; ++bb) {
      const size_t s = dg::sym4(aa, bb);
      a.out_g[s * n + p] = out.g[aa][bb];
      a.out_pi[s * n + p] = out.pi[aa][bb];
      for (int i = 0; i < 3; ++i) a.out_phi[(i + 3 * s) * n + p] = out.phi[i][aa][bb];
    }
}

extern "C" {

int dgrhs_gh_bjorhus_dg_time_derivative(
    int n, int physical, const double* normal_covector, const double* spacetime_metric,
    const double* pi, const double* phi, const double* coords, const double* gamma1,
    const double* gamma2, const double* lapse, const double* shift,
    const double* inverse_spacetime_metric, const double* spacetime_unit_normal_vector,
    const double* three_index_constraint, const double* gauge_source,
    const double* spacetime_deriv_gauge_source, const double* dt_spacetime_metric,
    const double* dt_pi, const double* dt_phi, const double* d_pi, const double* d_phi,
    double* dt_spacetime_metric_correction, double* dt_pi_correction,
    double* dt_phi_correction) {
  if (need_gpu()) return 1;
  if (n < 1) return fail("n must be positive");
  Staged st;
  BjorhusOpArgs a;
  a.n_lo = st.in(normal_covector, (size_t)3 * n);
  a.g = st.in(spacetime_metric, (size_t)10 * n);
  a.pi = st.in(pi, (size_t)10 * n);
  a.phi = st.in(phi, (size_t)30 * n);
  a.x = st.in(coords, (size_t)3 * n);
  a.gamma1 = st.in(gamma1, n);
  a.gamma2 = st.in(gamma2, n);
  a.lapse = st.in(lapse, n);
  a.shift = st.in(shift, (size_t)3 * n);
  a.ipsi = st.in(inverse_spacetime_metric, (size_t)10 * n);
  a.t_up = st.in(spacetime_unit_normal_vector, (size_t)4 * n);
  a.c3 = st.in(three_index_constraint, (size_t)30 * n);
  a.H = st.in(gauge_source, (size_t)4 * n);
  a.dH = st.in(spacetime_deriv_gauge_source, (size_t)16 * n);
  a.dt_g = st.in(dt_spacetime_metric, (size_t)10 * n);
  a.dt_pi = st.in(dt_pi, (size_t)10 * n);
  a.dt_phi = st.in(dt_phi, (size_t)30 * n);
  a.d_pi = st.in(d_pi, (size_t)30 * n);
  a.d_phi = st.in(d_phi, (size_t)90 * n);
  a.out_g = st.in(nullptr, (size_t)10 * n);
  a.out_pi = st.in(nullptr, (size_t)10 * n);
  a.out_phi = st.in(nullptr, (size_t)30 * n);
  if (!a.n_lo || !a.g || !a.pi || !a.phi || !a.x || !a.gamma1 || !a.gamma2 || !a.lapse ||
      !a.shift || !a.ipsi || !a.t_up || !a.c3 || !a.H || !a.dH || !a.dt_g || !a.dt_pi ||
      !a.dt_phi || !a.d_pi || !a.d_phi || !a.out_g || !a.out_pi || !a.out_phi)
    return fail("device staging failed");
  gh_bjorhus_op_kernel<<<grid(n), 128>>>(n, physical, a);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(dt_spacetime_metric_correction, a.out_g, (size_t)10 * n * 8,
                cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(dt_pi_correction, a.out_pi, (size_t)10 * n * 8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(dt_phi_correction, a.out_phi, (size_t)30 * n * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int dgrhs_gh_time_derivative(int n, const double* u, const double* du,
                             const double* gamma0, const double* gamma1,
                             const double* gamma2, int harmonic, const double* gauge_h,
                             const double* d4_gauge_h, double* dt_u) {
  if (need_gpu()) return 1;
  if (n < 1) return fail("n must be positive");
  if (!harmonic && (!gauge_h || !d4_gauge_h)) return fail("gauge fields required");
  Staged st;
  double* d_u = st.in(u, (size_t)50 * n);
  double* d_du = st.in(du, (size_t)150 * n);
  double* d_g0 = st.in(gamma0, n);
  double* d_g1 = st.in(gamma1, n);
  double* d_g2 = st.in(gamma2, n);
  double* d_H = st.in(harmonic ? nullptr : gauge_h, (size_t)4 * n);
  double* d_dH = st.in(harmonic ? nullptr : d4_gauge_h, (size_t)16 * n);
  double* d_dt = st.in(nullptr, (size_t)50 * n);
  if (!d_u || !d_du || !d_g0 || !d_g1 || !d_g2 || !d_H || !d_dH || !d_dt)
    return fail("device staging failed");
  if (harmonic)
    gh_time_derivative_kernel<true><<<grid(n), 128>>>(n, d_u, d_du, d_g0, d_g1, d_g2, d_H, d_dH, d_dt);
  else
    gh_time_derivative_kernel<false><<<grid(n), 128>>>(n, d_u, d_du, d_g0, d_g1, d_g2, d_H, d_dH, d_dt);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(dt_u, d_dt, (size_t)50 * n * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int dgrhs_sw_time_derivative(int n, const double* u, const double* du,
                             const double* gamma2, double* dt_u) {
  if (need_gpu()) return 1;
  if (n < 1) return fail("n must be positive");
  Staged st;
  double* d_u = st.in(u, (size_t)5 * n);
  double* d_du = st.in(du, (size_t)15 * n);
  double* d_g2 = st.in(gamma2, n);
  double* d_dt = st.in(nullptr, (size_t)5 * n);
  if (!d_u || !d_du || !d_g2 || !d_dt) return fail("device staging failed");
  sw_time_derivative_kernel<<<grid(n), 128>>>(n, d_u, d_du, d_g2, d_dt);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(dt_u, d_dt, (size_t)5 * n * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int dgrhs_gh_package_data_moving(int f, const double* u, const double* gamma1,
                                 const double* gamma2, const double* lapse, const double* shift,
                                 const double* normal_covector, const double* normal_vector,
                                 const double* normal_dot_mesh_velocity, double* packaged,
                                 double* max_abs_char_speed) {
  if (need_gpu()) return 1;
  if (f < 1) return fail("f must be positive");
  Staged st;
  double* d_u = st.in(u, (size_t)50 * f);
  double* d_g1 = st.in(gamma1, f);
  double* d_g2 = st.in(gamma2, f);
  double* d_l = st.in(lapse, f);
  double* d_s = st.in(shift, (size_t)3 * f);
  double* d_nl = st.in(normal_covector, (size_t)3 * f);
  double* d_nu = st.in(normal_vector, (size_t)3 * f);
  double* d_nv = normal_dot_mesh_velocity ? st.in(normal_dot_mesh_velocity, f) : nullptr;
  if (normal_dot_mesh_velocity && !d_nv) return fail("device staging failed");
  double* d_pk = st.in(nullptr, (size_t)134 * f);
  double* d_ms = st.in(nullptr, f);
  if (!d_u || !d_g1 || !d_g2 || !d_l || !d_s || !d_nl || !d_nu || !d_pk || !d_ms)
    return fail("device staging failed");
  gh_package_kernel<<<grid(f), 128>>>(f, d_u, d_g1, d_g2, d_l, d_s, d_nl, d_nu, d_nv, d_pk,
                                      d_ms);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(packaged, d_pk, (size_t)134 * f * 8, cudaMemcpyDeviceToHost));
  if (max_abs_char_speed) {
    std::vector<double> ms(f);
    CU(cudaMemcpy(ms.data(), d_ms, (size_t)f * 8, cudaMemcpyDeviceToHost));
    double m = ms[0];
    for (double v : ms) m = v > m ? v : m;
    *max_abs_char_speed = m;
  }
  return 0;
}

int dgrhs_gh_package_data(int f, const double* u, const double* gamma1,
                          const double* gamma2, const double* lapse, const double* shift,
                          const double* normal_covector, const double* normal_vector,
                          double* packaged, double* max_abs_char_speed) {
  return dgrhs_gh_package_data_moving(f, u, gamma1, gamma2, lapse, shift, normal_covector,
                                      normal_vector, nullptr, packaged, max_abs_char_speed);
}

int dgrhs_gh_boundary_terms(int f, const double* packaged_int, const double* packaged_ext,
                            double* boundary_correction) {
  if (need_gpu()) return 1;
  Staged st;
  double* d_i = st.in(packaged_int, (size_t)134 * f);
  double* d_e = st.in(packaged_ext, (size_t)134 * f);
  double* d_c = st.in(nullptr, (size_t)50 * f);
  if (!d_i || !d_e || !d_c) return fail("device staging failed");
  gh_boundary_terms_kernel<<<grid(f), 128>>>(f, d_i, d_e, d_c);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(boundary_correction, d_c, (size_t)50 * f * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int dgrhs_sw_package_data_moving(int f, const double* u, const double* gamma2,
                                 const double* normal_covector,
                                 const double* normal_dot_mesh_velocity, double* packaged,
                                 double* max_abs_char_speed) {
  if (need_gpu()) return 1;
  Staged st;
  double* d_u = st.in(u, (size_t)5 * f);
  double* d_g2 = st.in(gamma2, f);
  double* d_n = st.in(normal_covector, (size_t)3 * f);
  double* d_nv = normal_dot_mesh_velocity ? st.in(normal_dot_mesh_velocity, f) : nullptr;
  if (normal_dot_mesh_velocity && !d_nv) return fail("device staging failed");
  double* d_pk = st.in(nullptr, (size_t)16 * f);
  double* d_ms = st.in(nullptr, f);
  if (!d_u || !d_g2 || !d_n || !d_pk || !d_ms) return fail("device staging failed");
  sw_package_kernel<<<grid(f), 128>>>(f, d_u, d_g2, d_n, d_nv, d_pk, d_ms);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(packaged, d_pk, (size_t)16 * f * 8, cudaMemcpyDeviceToHost));
  if (max_abs_char_speed) {
    std::vector<double> ms(f);
    CU(cudaMemcpy(ms.data(), d_ms, (size_t)f * 8, cudaMemcpyDeviceToHost));
    double m = ms[0];
    for (double v : ms) m = v > m ? v : m;
    *max_abs_char_speed = m;
  }
  return 0;
}

int dgrhs_sw_package_data(int f, const double* u, const double* gamma2,
                          const double* normal_covector, double* packaged,
                          double* max_abs_char_speed) {
  return dgrhs_sw_package_data_moving(f, u, gamma2, normal_covector, nullptr, packaged,
                                      max_abs_char_speed);
}

int dgrhs_sw_boundary_terms(int f, const double* packaged_int, const double* packaged_ext,
                            double* boundary_correction) {
  if (need_gpu()) return 1;
  Staged st;
  double* d_i = st.in(packaged_int, (size_t)16 * f);
  double* d_e = st.in(packaged_ext, (size_t)16 * f);
  double* d_c = st.in(nullptr, (size_t)5 * f);
  if (!d_i || !d_e || !d_c) return fail("device staging failed");
  sw_boundary_terms_kernel<<<grid(f), 128>>>(f, d_i, d_e, d_c);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(boundary_correction, d_c, (size_t)5 * f * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int dgrhs_lift_flux(int f, int n_comps, double* boundary_correction,
                    int extent_perpendicular_to_boundary,
                    const double* magnitude_of_face_normal) {
  if (need_gpu()) return 1;
  Staged st;
  double* d_c = st.in(boundary_correction, (size_t)n_comps * f);
  double* d_m = st.in(magnitude_of_face_normal, f);
  if (!d_c || !d_m) return fail("device staging failed");
  lift_flux_kernel<<<grid(f), 128>>>(f, n_comps, d_c, extent_perpendicular_to_boundary, d_m);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(boundary_correction, d_c, (size_t)n_comps * f * 8, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"

// ---- dg::project_to_mortar / project_from_mortar, orient_variables_on_slice,
// ---- TimeStepper::update_u at operator granularity ---------------------------------

namespace {
// one pass of apply_matrices along one face dimension: out[c][..] = sum_s M[t][s] in[c][..]
// dims: in [n_comps][nb][na] (a fastest); along = 0: a, 1: b; M row-major [n_out][n_in]
__global__ void apply_matrix_2d_kernel(int n_comps, int na, int nb, int along, int n_out,
                                       const double* __restrict__ M,
                                       const double* __restrict__ in, double* __restrict__ out) {
  const int na_out = along == 0 ? n_out : na, nb_out = along == 1 ? n_out : nb;
  const int per = na_out * nb_out;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_comps * per) return;
  const int c = idx / per, p = idx % per, a = p % na_out, b = p / na_out;
  const double* src = in + (size_t)c * na * nb;
  double s = 0.0;
  if (along == 0) {
    for (int m = 0; m < na; ++m) s = fma(M[a * na + m], src[m + na * b], s);
  } else {
    for (int m = 0; m < nb; ++m) s = fma(M[b * nb + m], src[a + na * m], s);
  }
  out[idx] = s;
}

__global__ void orient_slice_kernel(int n_comps, int na, int nb, int perm,
                                    const double* __restrict__ in, double* __restrict__ out) {
  const int per = na * nb;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_comps * per) return;
  const int c = idx / per, p = idx % per, qa = p % na, qb = p / na;
  // the neighbour's extents and the point (qa, qb) in the neighbour's frame
  const int ma = (perm & 1) ? nb : na;
  int ta = (perm & 1) ? qb : qa, tb = (perm & 1) ? qa : qb;
  const int mb = (perm & 1) ? na : nb;
  if (perm & 2) ta = ma - 1 - ta;
  if (perm & 4) tb = mb - 1 - tb;
  out[(size_t)c * per + ta + ma * tb] = in[idx];
}

struct UpdateTerms {
  int n;
  double a;          // factor of u itself
  double c[9];
  const double* v[9];
};
__global__ void update_u_kernel(long long size, double* __restrict__ u, UpdateTerms t) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= size) return;
  double r = u[idx] * t.a;
  for (int j = 0; j < t.n; ++j) r = fma(t.c[j], t.v[j][idx], r);
  u[idx] = r;
}

// project along both face dimensions; to_mortar: face -> mortar (interpolation),
// else mortar -> face (L2 projection)
int project_face(bool to_mortar, int n_comps, const int* face_extents, const int* mortar_extents,
                 const int* mortar_size, const double* in, double* out) {
  if (need_gpu()) return 1;
  if (n_comps < 1) return fail("n_comps must be positive");
  for (int d = 0; d < 2; ++d) {
    if (face_extents[d] < 2 || mortar_extents[d] > 12 || mortar_extents[d] < face_extents[d])
      return fail("need 2 <= face extent <= mortar extent <= 12 (MortarHelpers.cpp:22-49: the "
                  "mortar mesh has the larger extents)");
    if (mortar_size[d] < 0 || mortar_size[d] > 2) return fail("bad mortar size");
  }
  Staged st;
  int ext[2] = {to_mortar ? face_extents[0] : mortar_extents[0],
                to_mortar ? face_extents[1] : mortar_extents[1]};
  const size_t n_in = (size_t)n_comps * ext[0] * ext[1];
  double* cur = st.in(in, n_in);
  if (!cur) return fail("device staging failed");
  for (int d = 0; d < 2; ++d) {
    const int n_face = face_extents[d], n_mortar = mortar_extents[d];
    // apply_matrices skips a dimension whose matrix is the identity (MortarHelpers.hpp:74-129)
    if (n_face == n_mortar && mortar_size[d] == 0) continue;
    const int n_out = to_mortar ? n_mortar : n_face, n_src = to_mortar ? n_face : n_mortar;
    std::vector<double> M((size_t)n_out * n_src);
    if (dgrhs_projection_matrix_meshes(n_face, n_mortar, to_mortar ? 0 : 1, mortar_size[d], M.data()))
      return 1;
    double* dM = st.in(M.data(), M.size());
    int nxt[2] = {ext[0], ext[1]};
    nxt[d] = n_out;
    double* dst = st.in(nullptr, (size_t)n_comps * nxt[0] * nxt[1]);
    if (!dM || !dst) return fail("device staging failed");
    const int total = n_comps * nxt[0] * nxt[1];
    apply_matrix_2d_kernel<<<(total + 127) / 128, 128>>>(n_comps, ext[0], ext[1], d, n_out, dM, cur, dst);
    dgrhs_internal_count_launch();
    CU(cudaGetLastError());
    cur = dst;
    ext[0] = nxt[0];
    ext[1] = nxt[1];
  }
  CU(cudaMemcpy(out, cur, (size_t)n_comps * ext[0] * ext[1] * 8, cudaMemcpyDeviceToHost));
  return 0;
}
}  // namespace

extern "C" {

int dgrhs_project_to_mortar(int n_comps, const int* face_extents, const int* mortar_extents,
                            const int* mortar_size, const double* face_vars, double* mortar_vars) {
  return project_face(true, n_comps, face_extents, mortar_extents, mortar_size, face_vars,
                      mortar_vars);
}

int dgrhs_project_from_mortar(int n_comps, const int* face_extents, const int* mortar_extents,
                              const int* mortar_size, const double* mortar_vars,
                              double* face_vars) {
  if (face_extents[0] == mortar_extents[0] && face_extents[1] == mortar_extents[1] &&
      mortar_size[0] == 0 && mortar_size[1] == 0)
    return fail("no projection is needed for a mortar that matches the face "
                "(MortarHelpers.hpp: needs_projection)");
  return project_face(false, n_comps, face_extents, mortar_extents, mortar_size, mortar_vars,
                      face_vars);
}

int dgrhs_orient_variables_on_slice(int n_comps, const int* slice_extents, int permutation,
                                    const double* vars, double* oriented) {
  if (need_gpu()) return 1;
  if (n_comps < 1 || slice_extents[0] < 1 || slice_extents[1] < 1 || permutation < 0 ||
      permutation > 7)
    return fail("bad arguments");
  Staged st;
  const size_t total = (size_t)n_comps * slice_extents[0] * slice_extents[1];
  double* d_in = st.in(vars, total);
  double* d_out = st.in(nullptr, total);
  if (!d_in || !d_out) return fail("device staging failed");
  orient_slice_kernel<<<(int)((total + 127) / 128), 128>>>(n_comps, slice_extents[0],
                                                            slice_extents[1], permutation, d_in,
                                                            d_out);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(oriented, d_out, total * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int dgrhs_update_u(int stepper, int order, long long size, double* u, int n_history,
                   const double* history_times, const double* history_derivatives,
                   const double* step_start_value, double time_step) {
  if (need_gpu()) return 1;
  if (size < 1 || n_history < 1 || n_history > 8) return fail("bad size / history length");
  UpdateTerms t{};
  t.a = 1.0;
  Staged st;
  double* d_u = st.in(u, (size_t)size);
  double* d_h = st.in(history_derivatives, (size_t)size * n_history);
  if (!d_u || !d_h) return fail("device staging failed");
  auto add = [&](double c, const double* v) {
    t.c[t.n] = c;
    t.v[t.n] = v;
    ++t.n;
  };
  if (stepper == DGRHS_STEPPER_ADAMS_BASHFORTH) {
    // AdamsBashforth::update_u_impl (AdamsBashforth.cpp:120-135): u += sum_j c_j f_j with the
    // coefficients of the history times (AdamsCoefficients.hpp:64-104), oldest term first
    if (order != n_history) return fail("Adams-Bashforth of order k needs k history entries");
    std::vector<double> coef(n_history);
    const double start = history_times[n_history - 1];
    if (dgrhs_adams_bashforth_coefficients(order, history_times, start, start + time_step,
                                           coef.data()))
      return 1;
    for (int j = 0; j < n_history; ++j) add(coef[j], d_h + (size_t)j * size);
  } else {
    int nsub = 0;
    if (dgrhs_stepper_properties(stepper, 0, nullptr, &nsub, nullptr, nullptr)) return 1;
    if (n_history > nsub) return fail("more history entries than substeps");
    if (!step_start_value) return fail("substep methods need the value at the start of the step");
    double* d_0 = st.in(step_start_value, (size_t)size);
    if (!d_0) return fail("device staging failed");
    const double* f_last = d_h + (size_t)(n_history - 1) * size;
    if (stepper == DGRHS_STEPPER_RK3_HESTHAVEN) {
      // Rk3HesthavenSsp.cpp:63-81; u holds the value of the current substep
      if (n_history == 1) {
        add(time_step, f_last);
      } else if (n_history == 2) {
        t.a = 0.25;
        add(0.75, d_0);
        add(0.25 * time_step, f_last);
      } else {
        t.a = 2.0 / 3.0;
        add(1.0 / 3.0, d_0);
        add((2.0 / 3.0) * time_step, f_last);
      }
    } else {
      // RungeKutta.cpp:69-122: u = u_start + dt sum_i coef_i f_i, the last substep with the
      // result coefficients
      std::vector<double> row(n_history);
      if (dgrhs_butcher_row(stepper, n_history - 1, row.data())) return 1;
      t.a = 0.0;
      add(1.0, d_0);
      for (int i = 0; i < n_history; ++i)
        if (row[i] != 0.0) add(row[i] * time_step, d_h + (size_t)i * size);
    }
  }
  update_u_kernel<<<(int)((size + 255) / 256), 256>>>(size, d_u, t);
  dgrhs_internal_count_launch();
  CU(cudaGetLastError());
  CU(cudaMemcpy(u, d_u, (size_t)size * 8, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
