// Pointwise physics of the DG right-hand side, shared by the volume and face
// kernels.  Everything here is per grid point, register resident and fully
// unrolled; nothing touches memory.  Functions are __host__ __device__ only so
// that tests/cpu_harness.cu can check the algebra against the oracle without a
// GPU -- the product path always runs them inside CUDA kernels.
//
// Reference (v2024.09.29) files restated here, with the algebra regrouped so
// that a thread needs few live registers (see DESIGN.md "GH volume kernel"):
//   Evolution/Systems/GeneralizedHarmonic/TimeDerivative.cpp:82-406
//   PointwiseFunctions/GeneralRelativity/{Shift,Lapse,InverseSpacetimeMetric,
//     Christoffel,SpacetimeNormalVector}.cpp
//   DataStructures/Tensor/EagerMath/DeterminantAndInverse.hpp:133-160
//   Evolution/Systems/GeneralizedHarmonic/BoundaryCorrections/UpwindPenalty.cpp
//   Evolution/Systems/ScalarWave/{TimeDerivative.cpp,BoundaryCorrections/
//     UpwindPenalty.cpp}
//   Evolution/DiscontinuousGalerkin/NormalCovectorAndMagnitude.hpp:47-92
#pragma once

#ifdef __CUDACC__
#define DG_HD __host__ __device__ __forceinline__
#else
#define DG_HD inline
#endif

namespace dg {

// storage index of the symmetric pair (a, b), Tensor/Structure.hpp:162-194
DG_HD constexpr int sym4(int a, int b) {
  return a <= b ? a * 4 - a * (a - 1) / 2 + (b - a)
                : b * 4 - b * (b - 1) / 2 + (a - b);
}
DG_HD constexpr int sym3(int a, int b) {
  return a <= b ? a * 3 - a * (a - 1) / 2 + (b - a)
                : b * 3 - b * (b - 1) / 2 + (a - b);
}

// ---------------------------------------------------------------------------
// 3+1 quantities from the spacetime metric (10 independent components)
// ---------------------------------------------------------------------------
struct Geom3p1 {
  double lapse;
  double shift[3];
  double ig[6];   // inverse spatial metric gamma^{ij}, sym3 order
  double det;     // det gamma_ij
};

DG_HD void geom_from_metric(const double (&g)[10], Geom3p1& q) {
  const double t00 = g[sym4(1, 1)], t01 = g[sym4(1, 2)], t02 = g[sym4(1, 3)];
  const double t11 = g[sym4(2, 2)], t12 = g[sym4(2, 3)], t22 = g[sym4(3, 3)];
  const double a = t11 * t22 - t12 * t12;
  const double b = t12 * t02 - t01 * t22;
  const double c = t01 * t12 - t11 * t02;
  q.det = t00 * a + t01 * b + t02 * c;
  const double inv = 1.0 / q.det;
  q.ig[0] = a * inv;
  q.ig[1] = b * inv;
  q.ig[2] = c * inv;
  q.ig[3] = (t22 * t00 - t02 * t02) * inv;
  q.ig[4] = (t02 * t01 - t00 * t12) * inv;
  q.ig[5] = (t00 * t11 - t01 * t01) * inv;
  double l = -g[0];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double s = q.ig[sym3(i, 0)] * g[sym4(1, 0)];
    s += q.ig[sym3(i, 1)] * g[sym4(2, 0)];
    s += q.ig[sym3(i, 2)] * g[sym4(3, 0)];
    q.shift[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) l += q.shift[i] * g[sym4(i + 1, 0)];
  q.lapse = sqrt(l);
}

// ---------------------------------------------------------------------------
// GH volume context: everything the per-(mu,nu) streaming phase needs.
// ---------------------------------------------------------------------------
struct GhContext {
  double lapse;
  double shift[3];      // beta^m (inertial)
  double shift_hat[3];  // J(jhat, m) beta^m : shift in the logical frame
  double gamma1, gamma2;
  double half_pi_nn;    // 1/2 n^a n^b Pi_ab
  double w[3];          // (n^a Pi_a,k+1) gamma^{km}
  double Gj[3][3];      // Gj[jhat][n] = J(jhat, m) gamma^{mn}
  double J[3][3];       // J[jhat][i]
  double half_phi_nn[3];
  double V[3][3];       // V[i][m] = (n^a Phi_i,a,n+1) gamma^{nm}
};

// gauge source function at a point: H_a and d_a H_b (dH[a][b])
struct GaugeH {
  double H[4];
  double dH[4][4];
};

// DampedHarmonic gauge parameters (GaugeSourceFunctions/DampedHarmonic.hpp)
struct DampedHarmonicParams {
  double sigma_r;
  double amp_L1, amp_L2, amp_S;
  int exp_L1, exp_L2, exp_S;
};

DG_HD double integer_pow(double x, int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= x;
  return r;
}

// damped_harmonic_impl<UseRollon = false> (DampedHarmonic.cpp:70-439,
// DampedWaveHelpers.cpp:26-62) at one point.  dag[a][sym4(b,c)] = d_a g_bc.
DG_HD void damped_harmonic_gauge(const DampedHarmonicParams& prm, const double (&x)[3],
                                 double lapse, const double (&shift)[3], double sqrt_det,
                                 const double (&ig)[6], const double (&dag)[4][10],
                                 double half_pi_nn, const double (&half_phi_nn)[3],
                                 const double (&g)[10], GaugeH& out) {
  const double one_over_lapse = 1.0 / lapse;
  const double log_fac_1 = log(sqrt_det / lapse);
  const double log_fac_2 = -log(lapse);
  const double inv_s2 = 1.0 / (prm.sigma_r * prm.sigma_r);
  const double weight = exp(-(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) * inv_s2);
  double d4_weight[4];
  d4_weight[0] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) d4_weight[i + 1] = -2.0 * weight * inv_s2 * x[i];
  double pow1 = integer_pow(log_fac_1, prm.exp_L1);
  double pow2 = integer_pow(log_fac_1, prm.exp_S);
  double pow3 = integer_pow(log_fac_2, prm.exp_L2);
  const double mu_L1 = prm.amp_L1 * weight * pow1;
  const double mu_S = prm.amp_S * weight * pow2;
  const double mu_L2 = prm.amp_L2 * weight * pow3;
  const double mu_S_over_lapse = mu_S * one_over_lapse;
  const double mu1 = mu_L1 * log_fac_1, mu2 = mu_L2 * log_fac_2;
  const double prefac = mu1 + mu2;
  double g_dot_shift[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double v = g[sym4(a, 1)] * shift[0];
    v += g[sym4(a, 2)] * shift[1];
    v += g[sym4(a, 3)] * shift[2];
    g_dot_shift[a] = v;
    out.H[a] = -mu_S_over_lapse * v;
  }
  out.H[0] -= prefac * lapse;
  double sh_hphi = shift[0] * half_phi_nn[0];
  sh_hphi += shift[1] * half_phi_nn[1];
  sh_hphi += shift[2] * half_phi_nn[2];
  const double dt_lapse = lapse * (lapse * half_pi_nn - sh_hphi);
  double dlbl[4];
  dlbl[0] = one_over_lapse * dt_lapse;
#pragma unroll
  for (int i = 0; i < 3; ++i) dlbl[i + 1] = -half_phi_nn[i];
  const double c1 = (double)(prm.exp_L1 + 1) * integer_pow(log_fac_1, prm.exp_L1);
  const double cS = (double)prm.exp_S * integer_pow(log_fac_1, prm.exp_S - 1);
  const double c2 = (double)(prm.exp_L2 + 1) * integer_pow(log_fac_2, prm.exp_L2);
  pow1 *= log_fac_1 * prm.amp_L1;
  pow2 *= prm.amp_S;
  pow3 *= log_fac_2 * prm.amp_L2;
  double d4_mu12[4], d4_mu_S[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double dgd = 0.0;  // gamma^{jk} d_a gamma_jk
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = j; k < 3; ++k)
        dgd += (j == k ? 1.0 : 2.0) * ig[sym3(j, k)] * dag[a][sym4(j + 1, k + 1)];
    const double d_logfac_1 = 0.5 * dgd - dlbl[a];
    const double d_logfac_2 = -dlbl[a];
    const double d4_mu1 = pow1 * d4_weight[a] + prm.amp_L1 * weight * (c1 * d_logfac_1);
    const double d4_mu2 = pow3 * d4_weight[a] + prm.amp_L2 * weight * (c2 * d_logfac_2);
    d4_mu12[a] = d4_mu1 + d4_mu2;
    d4_mu_S[a] = d4_weight[a] * pow2 + prm.amp_S * weight * (cS * d_logfac_1);
  }
  double d4_muS_ol[4], dT2[4];
  dT2[0] = -d4_mu12[0] * lapse - prefac * dt_lapse;
  d4_muS_ol[0] = (dt_lapse * (-mu_S * one_over_lapse) + d4_mu_S[0]) * one_over_lapse;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dT2[i + 1] = -d4_mu12[i + 1] * lapse + prefac * lapse * half_phi_nn[i];
    d4_muS_ol[i + 1] = one_over_lapse * (d4_mu_S[i + 1] + mu_S * half_phi_nn[i]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double dT3[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) dT3[j + 1] = dag[a][sym4(0, j + 1)];
    double v = dag[a][sym4(0, 1)] * shift[0];
    v += dag[a][sym4(0, 2)] * shift[1];
    v += dag[a][sym4(0, 3)] * shift[2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i + 1; j < 3; ++j) v -= shift[i] * shift[j] * dag[a][sym4(i + 1, j + 1)];
    v *= 2.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) v -= shift[i] * shift[i] * dag[a][sym4(i + 1, i + 1)];
    dT3[0] = v;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      out.dH[a][b] = dT3[b] * (-mu_S_over_lapse) - d4_muS_ol[a] * g_dot_shift[b];
    out.dH[a][0] += dT2[a];
  }
}

// what the gauge condition needs at a point
struct GaugeInput {
  const GaugeH* fields;          // kGauge == 1
  DampedHarmonicParams dh;       // kGauge == 2
  double x[3];                   // kGauge == 2: inertial coordinates
  GaugeH* computed = nullptr;    // kGauge == 2: if set, receives H_a and d_a H_b
  // kGauge == 1, device only: if set, H_a (component a at H_global[a * stride]) and d_a H_b
  // (component a + 4 b at dH_global[(a + 4 b) * stride]) are fetched inside the prologue just
  // before their first use instead of being handed over in `fields`: twenty values fewer in
  // registers while the 50 evolved components arrive and the normal contractions are formed
  const double* H_global = nullptr;
  const double* dH_global = nullptr;
  size_t stride = 0;
};

// Computes the context and Q[10], the part of the bracket of the dt Pi
// equation that contains no derivatives and is not linear in the pair's own
// components (TimeDerivative.cpp:308-372): constraint-damping n_a terms, the
// three quadratic contractions and the gauge terms.
//   g, pi: sym4 order; phi[m][sym4]; J[jhat][i]
// kGauge: 0 Harmonic, 1 fields supplied, 2 DampedHarmonic
// Folds the inverse Jacobian into the linear coefficients of the context
// (shift_hat, Gj, J); ig = gamma^{ij} as returned by gh_prologue_core.
DG_HD void gh_context_set_jacobian(GhContext& ctx, const double (&J)[3][3],
                                   const double (&ig)[6]) {
#pragma unroll
  for (int jh = 0; jh < 3; ++jh) {
    double v = J[jh][0] * ctx.shift[0];
    v += J[jh][1] * ctx.shift[1];
    v += J[jh][2] * ctx.shift[2];
    ctx.shift_hat[jh] = v;
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      double s = J[jh][0] * ig[sym3(0, n)];
      s += J[jh][1] * ig[sym3(1, n)];
      s += J[jh][2] * ig[sym3(2, n)];
      ctx.Gj[jh][n] = s;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) ctx.J[jh][i] = J[jh][i];
  }
}

// Q may be an array of 10 doubles or any object with operator[] that returns a reference
// (QStrided: straight into shared memory, so that the ten accumulators do not hold registers
// through the register-critical middle of the prologue)
// Orders the shared-memory updates of Q (and with them the arithmetic that feeds them): keeps
// the compiler from interleaving the unrolled iterations of the quadratic terms, whose
// temporaries would otherwise all be live at once.  No instruction is emitted.
#if defined(__CUDA_ARCH__) && !defined(DG_NO_SCHED_BARRIER)
#define DG_SCHED_BARRIER() asm volatile("" ::: "memory")
#else
#define DG_SCHED_BARRIER() ((void)0)
#endif
struct QStrided {
  double* p;
  int stride;
  DG_HD double& operator[](int s) const { return p[s * stride]; }
};
// Sixteen values that are formed early and needed only by the streaming phase (w, V,
// half_pi_nn, half_phi_nn) are handed to a "park" object between the two: LocalPark keeps them
// in registers (host code, the small kernels), the fused volume kernel passes shared memory
// that is idle during the prologue, so that they do not occupy registers through the
// quadratic terms.
struct LocalPark {
  double v[16];
  DG_HD double& operator[](int k) { return v[k]; }
};
template <int kGauge, typename QRef, typename Park>
DG_HD void gh_prologue_core(const double (&g)[10], const double (&pi)[10],
                            const double (&phi)[3][10], double gamma0, double gamma1,
                            double gamma2, const GaugeInput& gin, GhContext& ctx,
                            QRef&& Q, double (&ig_out)[6], Park&& park);
template <int kGauge, typename QRef>
DG_HD void gh_prologue_core(const double (&g)[10], const double (&pi)[10],
                            const double (&phi)[3][10], double gamma0, double gamma1,
                            double gamma2, const GaugeInput& gin, GhContext& ctx,
                            QRef&& Q, double (&ig_out)[6]) {
  gh_prologue_core<kGauge>(g, pi, phi, gamma0, gamma1, gamma2, gin, ctx, Q, ig_out, LocalPark{});
}

template <int kGauge>
DG_HD void gh_prologue(const double (&g)[10], const double (&pi)[10],
                       const double (&phi)[3][10], const double (&J)[3][3],
                       double gamma0, double gamma1, double gamma2,
                       const GaugeInput& gin, GhContext& ctx, double (&Q)[10]) {
  double ig[6];
  gh_prologue_core<kGauge>(g, pi, phi, gamma0, gamma1, gamma2, gin, ctx, Q, ig);
  gh_context_set_jacobian(ctx, J, ig);
}

template <int kGauge, typename QRef, typename Park>
DG_HD void gh_prologue_core(const double (&g)[10], const double (&pi)[10],
                            const double (&phi)[3][10], double gamma0, double gamma1,
                            double gamma2, const GaugeInput& gin, GhContext& ctx,
                            QRef&& Q, double (&ig_out)[6], Park&& park) {
  constexpr bool kHarmonic = kGauge == 0;
  Geom3p1 q;
  geom_from_metric(g, q);
  const double lapse = q.lapse;
  ctx.lapse = lapse;
  ctx.gamma1 = gamma1;
  ctx.gamma2 = gamma2;
  // inverse spacetime metric (InverseSpacetimeMetric.cpp:27-48), sym4 order
  double G[10];
  const double m1ol2 = -1.0 / (lapse * lapse);
  G[0] = m1ol2;
#pragma unroll
  for (int i = 0; i < 3; ++i) G[sym4(0, i + 1)] = -q.shift[i] * m1ol2;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      G[sym4(i + 1, j + 1)] = q.ig[sym3(i, j)] + q.shift[i] * q.shift[j] * m1ol2;
  // unit normal vector (SpacetimeNormalVector.cpp:28-38)
  double nv[4];
  nv[0] = 1.0 / lapse;
#pragma unroll
  for (int i = 0; i < 3; ++i) nv[i + 1] = -q.shift[i] * nv[0];

  // d_a g_bc: a = 0 -> -lapse Pi + shift^m Phi_m, a = m+1 -> Phi_m
  double dag[4][10];
#pragma unroll
  for (int s = 0; s < 10; ++s) {
    double v = -lapse * pi[s];
#pragma unroll
    for (int m = 0; m < 3; ++m) v += q.shift[m] * phi[m][s];
    dag[0][s] = v;
#pragma unroll
    for (int m = 0; m < 3; ++m) dag[m + 1][s] = phi[m][s];
  }
  // normal contractions (TimeDerivative.cpp:192-231)
  double pon[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double v = nv[0] * pi[sym4(0, a)];
#pragma unroll
    for (int b = 1; b < 4; ++b) v += nv[b] * pi[sym4(b, a)];
    pon[a] = v;
  }
  {
    double v = nv[0] * pon[0];
#pragma unroll
    for (int a = 1; a < 4; ++a) v += nv[a] * pon[a];
    if constexpr (kGauge == 2) ctx.half_pi_nn = 0.5 * v; else park[12] = 0.5 * v;
  }
  double pho[3][4];
#pragma unroll
  for (int n = 0; n < 3; ++n) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double v = nv[0] * phi[n][sym4(0, a)];
#pragma unroll
      for (int b = 1; b < 4; ++b) v += nv[b] * phi[n][sym4(b, a)];
      pho[n][a] = v;
    }
    double v = nv[0] * pho[n][0];
#pragma unroll
    for (int a = 1; a < 4; ++a) v += nv[a] * pho[n][a];
    if constexpr (kGauge == 2) ctx.half_phi_nn[n] = 0.5 * v; else park[13 + n] = 0.5 * v;
  }
  // w and V of the linear-coefficient context need only these contractions and gamma^{ij}:
  // formed now and parked (see LocalPark)
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    double v = pon[1] * q.ig[sym3(0, m)];
    v += pon[2] * q.ig[sym3(1, m)];
    v += pon[3] * q.ig[sym3(2, m)];
    park[m] = v;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      double v = pho[i][1] * q.ig[sym3(0, m)];
      v += pho[i][2] * q.ig[sym3(1, m)];
      v += pho[i][3] * q.ig[sym3(2, m)];
      park[3 + 3 * i + m] = v;
    }
  // Christoffel symbols of the first kind Gamma_k,ij (Christoffel.cpp:16-31)
  // are formed on the fly: chr(k,i,j) = 1/2 (d_i g_jk + d_j g_ik - d_k g_ij)
#define DG_CHR(k, i, j) \
  (0.5 * (dag[i][sym4(j, k)] + dag[j][sym4(i, k)] - dag[k][sym4(i, j)]))
  GaugeH gauge_local;
  const GaugeH* gauge = gin.fields;
#ifdef __CUDA_ARCH__
  if constexpr (kGauge == 1) {
    if (gin.H_global != nullptr) {
      asm volatile("" ::: "memory");   // not before the work above
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        gauge_local.H[a] = __ldg(gin.H_global + (size_t)a * gin.stride);
#pragma unroll
        for (int b = 0; b < 4; ++b)
          gauge_local.dH[a][b] = __ldg(gin.dH_global + (size_t)(a + 4 * b) * gin.stride);
      }
      gauge = &gauge_local;
    }
  }
#endif
  if constexpr (kGauge == 2) {
    damped_harmonic_gauge(gin.dh, gin.x, lapse, q.shift, sqrt(q.det), q.ig, dag,
                          ctx.half_pi_nn, ctx.half_phi_nn, g, gauge_local);
    gauge = &gauge_local;
    if (gin.computed) *gin.computed = gauge_local;
  }
  // gauge constraint C_a = Gamma_a + H_a, Gamma_a = G^{bc} Gamma_a,bc
  double Ca[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double v = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = b; c < 4; ++c)
        v += (b == c ? 1.0 : 2.0) * DG_CHR(a, b, c) * G[sym4(b, c)];
    Ca[a] = v;
    if (!kHarmonic) Ca[a] += gauge->H[a];
  }
  double nC = nv[0] * Ca[0];
#pragma unroll
  for (int a = 1; a < 4; ++a) nC += nv[a] * Ca[a];
  nC *= gamma0;
  // n_a pieces (TimeDerivative.cpp:308-340)
  const double base = -gamma0 * lapse;
  Q[0] = 2.0 * base * Ca[0] - nC * g[0];
#pragma unroll
  for (int i = 1; i < 4; ++i) Q[sym4(0, i)] = base * Ca[i] - nC * g[sym4(0, i)];
#pragma unroll
  for (int i = 1; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) Q[sym4(i, j)] = -nC * g[sym4(i, j)];
  if (!kHarmonic) {
    // -(d_mu H_nu + d_nu H_mu) + 2 Gamma^d_{mu nu} H_d
    double Hup[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double v = 0.0;
#pragma unroll
      for (int d = 0; d < 4; ++d) v += G[sym4(e, d)] * gauge->H[d];
      Hup[e] = v;
    }
#pragma unroll
    for (int mu = 0; mu < 4; ++mu)
#pragma unroll
      for (int nu = mu; nu < 4; ++nu) {
        double v = -(gauge->dH[mu][nu] + gauge->dH[nu][mu]);
#pragma unroll
        for (int e = 0; e < 4; ++e) v += 2.0 * Hup[e] * DG_CHR(e, mu, nu);
        Q[sym4(mu, nu)] += v;
      }
  }
  // T1 = Pi G Pi  (:350 "2 pi(mu,delta) pi_2_up(nu,delta)")
  {
    double X[4][4];  // X[nu][d] = G^{d b} Pi_{nu b}
#pragma unroll
    for (int nu = 0; nu < 4; ++nu)
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        double v = G[sym4(d, 0)] * pi[sym4(nu, 0)];
#pragma unroll
        for (int b = 1; b < 4; ++b) v += G[sym4(d, b)] * pi[sym4(nu, b)];
        X[nu][d] = v;
      }
#pragma unroll
    for (int mu = 0; mu < 4; ++mu)
#pragma unroll
      for (int nu = mu; nu < 4; ++nu) {
        double v = 0.0;
#pragma unroll
        for (int d = 0; d < 4; ++d) v += pi[sym4(mu, d)] * X[nu][d];
        Q[sym4(mu, nu)] -= 2.0 * v;
      }
    DG_SCHED_BARRIER();
  }
  // T2 = sum_n (gamma^{nm} Phi_m) G Phi_n  (:355-358)
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    double Y[10];
#pragma unroll
    for (int s = 0; s < 10; ++s) {
      double v = q.ig[sym3(n, 0)] * phi[0][s];
      v += q.ig[sym3(n, 1)] * phi[1][s];
      v += q.ig[sym3(n, 2)] * phi[2][s];
      Y[s] = v;
    }
    double X[4][4];  // X[nu][d] = G^{d b} Phi_{n nu b}
#pragma unroll
    for (int nu = 0; nu < 4; ++nu)
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        double v = G[sym4(d, 0)] * phi[n][sym4(nu, 0)];
#pragma unroll
        for (int b = 1; b < 4; ++b) v += G[sym4(d, b)] * phi[n][sym4(nu, b)];
        X[nu][d] = v;
      }
#pragma unroll
    for (int mu = 0; mu < 4; ++mu)
#pragma unroll
      for (int nu = mu; nu < 4; ++nu) {
        double v = 0.0;
#pragma unroll
        for (int d = 0; d < 4; ++d) v += Y[sym4(mu, d)] * X[nu][d];
        Q[sym4(mu, nu)] += 2.0 * v;
      }
    DG_SCHED_BARRIER();
  }
  // T3_{mu nu} = tr(Gamma_mu G Gamma_nu G) = Gamma_nu,ab C_mu^{ab},
  // C_mu = G Gamma_mu G (symmetric)  (:360-364).  One C_mu is live at a time.
#pragma unroll
  for (int mu = 0; mu < 4; ++mu) {
    double C[10];
    {
      double A[4][4];  // A[a][d] = Gamma_mu,a,b G^{b d}
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          double v = DG_CHR(mu, a, 0) * G[sym4(0, d)];
#pragma unroll
          for (int b = 1; b < 4; ++b) v += DG_CHR(mu, a, b) * G[sym4(b, d)];
          A[a][d] = v;
        }
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int d = e; d < 4; ++d) {
          double v = G[sym4(e, 0)] * A[0][d];
#pragma unroll
          for (int a = 1; a < 4; ++a) v += G[sym4(e, a)] * A[a][d];
          C[sym4(e, d)] = (e == d) ? v : 2.0 * v;
        }
    }
#pragma unroll
    for (int nu = mu; nu < 4; ++nu) {
      double v = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a; b < 4; ++b) v += DG_CHR(nu, a, b) * C[sym4(a, b)];
      Q[sym4(mu, nu)] -= 2.0 * v;
    }
    DG_SCHED_BARRIER();
  }
#undef DG_CHR
  // linear-coefficient context
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    ctx.shift[m] = q.shift[m];
    ctx.w[m] = park[m];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int m = 0; m < 3; ++m) ctx.V[i][m] = park[3 + 3 * i + m];
  if constexpr (kGauge != 2) {
    ctx.half_pi_nn = park[12];
#pragma unroll
    for (int n = 0; n < 3; ++n) ctx.half_phi_nn[n] = park[13 + n];
  }
#pragma unroll
  for (int x = 0; x < 6; ++x) ig_out[x] = q.ig[x];
}

// Gamma_a = g^{bc} Gamma_a,bc from the evolved variables (the trace that
// gh::TimeDerivative forms at TimeDerivative.cpp:125-128); the gauge constraint
// is C_a = H_a + Gamma_a (Constraints.cpp:965-1000).
DG_HD void gh_trace_christoffel(const double (&g)[10], const double (&pi)[10],
                                const double (&phi)[3][10], double (&Gam)[4]) {
  Geom3p1 q;
  geom_from_metric(g, q);
  double G[10];
  const double m1ol2 = -1.0 / (q.lapse * q.lapse);
  G[0] = m1ol2;
#pragma unroll
  for (int i = 0; i < 3; ++i) G[sym4(0, i + 1)] = -q.shift[i] * m1ol2;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j)
      G[sym4(i + 1, j + 1)] = q.ig[sym3(i, j)] + q.shift[i] * q.shift[j] * m1ol2;
  double dag[4][10];
#pragma unroll
  for (int s = 0; s < 10; ++s) {
    double v = -q.lapse * pi[s];
#pragma unroll
    for (int m = 0; m < 3; ++m) v += q.shift[m] * phi[m][s];
    dag[0][s] = v;
#pragma unroll
    for (int m = 0; m < 3; ++m) dag[m + 1][s] = phi[m][s];
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double v = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = b; c < 4; ++c)
        v += (b == c ? 1.0 : 2.0) * G[sym4(b, c)] * 0.5 *
             (dag[b][sym4(c, a)] + dag[c][sym4(b, a)] - dag[a][sym4(b, c)]);
    Gam[a] = v;
  }
}

// One (mu,nu) pair: the pair's own five components, their 15 logical
// derivatives (d*[jhat]) and Q -> the five time derivatives
// (TimeDerivative.cpp:296-306, 342-392, 394-412).
DG_HD void gh_pair_rhs(const GhContext& c, double Qs, double g, double pi,
                       const double (&ph)[3], const double (&dg)[3],
                       const double (&dpi)[3], const double (&dph)[3][3],
                       double& out_g, double& out_pi, double (&out_phi)[3]) {
  double sphi = c.shift[0] * ph[0];
  sphi += c.shift[1] * ph[1];
  sphi += c.shift[2] * ph[2];
  double sg = c.shift_hat[0] * dg[0];
  sg += c.shift_hat[1] * dg[1];
  sg += c.shift_hat[2] * dg[2];
  const double c3s = sg - sphi;  // shift^m (d_m g - Phi_m)
  out_g = (-c.lapse * pi + sphi) + (1.0 + c.gamma1) * c3s;
  double t = Qs - c.half_pi_nn * pi;
#pragma unroll
  for (int m = 0; m < 3; ++m) t -= c.w[m] * ph[m];
#pragma unroll
  for (int jh = 0; jh < 3; ++jh)
#pragma unroll
    for (int n = 0; n < 3; ++n) t -= c.Gj[jh][n] * dph[n][jh];
  double o = c.lapse * t + (c.gamma1 * c.gamma2) * c3s;
#pragma unroll
  for (int jh = 0; jh < 3; ++jh) o += c.shift_hat[jh] * dpi[jh];
  out_pi = o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double dgi = c.J[0][i] * dg[0];
    dgi += c.J[1][i] * dg[1];
    dgi += c.J[2][i] * dg[2];
    double dpii = c.J[0][i] * dpi[0];
    dpii += c.J[1][i] * dpi[1];
    dpii += c.J[2][i] * dpi[2];
    double v = pi * c.half_phi_nn[i] - dpii + c.gamma2 * (dgi - ph[i]);
#pragma unroll
    for (int m = 0; m < 3; ++m) v += c.V[i][m] * ph[m];
    v *= c.lapse;
#pragma unroll
    for (int jh = 0; jh < 3; ++jh) v += c.shift_hat[jh] * dph[i][jh];
    out_phi[i] = v;
  }
}

// ---------------------------------------------------------------------------
// Two-kernel variant of the GH volume work (context kernel + streaming kernel):
// the context kernel stores, per point, the 26 values that need all 50
// components (Q, and the normal contractions raised with gamma^{ij}); the
// streaming kernel recomputes lapse/shift/gamma^{ij} from g and works with
// inertial derivatives.
// ---------------------------------------------------------------------------
constexpr int kGhCtxComps = 26;  // Q 10 | half_pi_nn | w 3 | half_phi_nn 3 | V 9

struct GhStreamCtx {
  double lapse, shift[3], ig[6];
  double gamma1, gamma2, half_pi_nn;
  double w[3], half_phi_nn[3];
};

// same equations as gh_pair_rhs, with inertial derivatives:
//   dgi[i] = d_i g, dpii[i] = d_i Pi, dphi[n][i] = d_i Phi_n; V[i][m]
DG_HD void gh_pair_rhs_inertial(const GhStreamCtx& c, const double (&V)[3][3], double Qs,
                                double g, double pi, const double (&ph)[3],
                                const double (&dgi)[3], const double (&dpii)[3],
                                const double (&dphi)[3][3], double& out_g, double& out_pi,
                                double (&out_phi)[3]) {
  double sphi = c.shift[0] * ph[0];
  sphi += c.shift[1] * ph[1];
  sphi += c.shift[2] * ph[2];
  double sg = c.shift[0] * dgi[0];
  sg += c.shift[1] * dgi[1];
  sg += c.shift[2] * dgi[2];
  const double c3s = sg - sphi;
  out_g = (-c.lapse * pi + sphi) + (1.0 + c.gamma1) * c3s;
  double t = Qs - c.half_pi_nn * pi;
#pragma unroll
  for (int m = 0; m < 3; ++m) t -= c.w[m] * ph[m];
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int n = 0; n < 3; ++n) t -= c.ig[sym3(m, n)] * dphi[n][m];
  double o = c.lapse * t + (c.gamma1 * c.gamma2) * c3s;
#pragma unroll
  for (int m = 0; m < 3; ++m) o += c.shift[m] * dpii[m];
  out_pi = o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double v = pi * c.half_phi_nn[i] - dpii[i] + c.gamma2 * (dgi[i] - ph[i]);
#pragma unroll
    for (int m = 0; m < 3; ++m) v += V[i][m] * ph[m];
    v *= c.lapse;
#pragma unroll
    for (int m = 0; m < 3; ++m) v += c.shift[m] * dphi[i][m];
    out_phi[i] = v;
  }
}

// ---------------------------------------------------------------------------
// Faces.  One side of an interface at one face point.
// ---------------------------------------------------------------------------
struct GhFaceSide {
  double n_lo[3], n_up[3];
  double mag;          // |n| of the unnormalised covector (for the lift)
  double gamma2;
  double speed[4];     // lambda_g, lambda_0, lambda_+, lambda_-
};

// unnorm = +-row `dim` of the inverse Jacobian on the face
// (InternalMortarDataImpl.hpp:180-221); curved normalisation with gamma^{ij}
// (NormalCovectorAndMagnitude.hpp:47-92); char speeds UpwindPenalty.cpp:76-88
DG_HD void gh_face_side(const double (&g)[10], const double (&unnorm)[3],
                        double gamma1, double gamma2, GhFaceSide& s) {
  Geom3p1 q;
  geom_from_metric(g, q);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double v = q.ig[sym3(i, 0)] * unnorm[0];
    v += q.ig[sym3(i, 1)] * unnorm[1];
    v += q.ig[sym3(i, 2)] * unnorm[2];
    s.n_up[i] = v;
  }
  double m = s.n_up[0] * unnorm[0];
  m += s.n_up[1] * unnorm[1];
  m += s.n_up[2] * unnorm[2];
  s.mag = sqrt(m);
  const double inv = 1.0 / s.mag;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    s.n_lo[i] = unnorm[i] * inv;
    s.n_up[i] *= inv;
  }
  double sdn = q.shift[0] * s.n_lo[0];
  sdn += q.shift[1] * s.n_lo[1];
  sdn += q.shift[2] * s.n_lo[2];
  sdn = -sdn;
  s.speed[1] = sdn;
  s.speed[0] = (1.0 + gamma1) * sdn;
  s.speed[2] = q.lapse + sdn;
  s.speed[3] = -q.lapse + sdn;
  s.gamma2 = gamma2;
}

// Moving mesh: the characteristic speeds relative to the mesh (dg_package_data with
// normal_dot_mesh_velocity, GH UpwindPenalty.cpp:85-91 / ScalarWave UpwindPenalty.cpp:55-67):
// v = inertial mesh velocity at the face point
DG_HD void face_side_mesh_velocity(GhFaceSide& s, const double (&v)[3], double one_plus_gamma1) {
  double ndv = s.n_lo[0] * v[0];
  ndv += s.n_lo[1] * v[1];
  ndv += s.n_lo[2] * v[2];
  s.speed[0] -= ndv * one_plus_gamma1;
  s.speed[1] -= ndv;
  s.speed[2] -= ndv;
  s.speed[3] -= ndv;
}

// characteristic-speed weighted fields of one (a,b) pair on one side
// (dg_package_data, UpwindPenalty.cpp:90-150)
struct GhPairPackaged {
  double v_g, g2_v_g, v_plus, v_minus, v_zero[3];
};

DG_HD void gh_pair_package(const GhFaceSide& s, double g, double pi,
                           const double (&ph)[3], GhPairPackaged& k) {
  const double g2g = s.gamma2 * g;
  double ndphi = s.n_up[0] * ph[0];
  ndphi += s.n_up[1] * ph[1];
  ndphi += s.n_up[2] * ph[2];
  k.v_plus = s.speed[2] * (pi + ndphi - g2g);
  k.v_minus = s.speed[3] * (pi - ndphi - g2g);
#pragma unroll
  for (int i = 0; i < 3; ++i) k.v_zero[i] = s.speed[1] * (ph[i] - s.n_lo[i] * ndphi);
  k.v_g = s.speed[0] * g;
  k.g2_v_g = g2g * s.speed[0];
}

DG_HD double step_function(double x) { return x < 0.0 ? 0.0 : 1.0; }

// dg_boundary_terms for one pair (UpwindPenalty.cpp:161-275) followed by the
// lift factor (LiftFlux.hpp:57-61) applied by the caller.
DG_HD void gh_pair_boundary_terms(const GhFaceSide& in, const GhFaceSide& ex,
                                  const GhPairPackaged& ki,
                                  const GhPairPackaged& ke, double& c_g,
                                  double& c_pi, double (&c_phi)[3]) {
  const double w_g_i = step_function(-in.speed[0]), w_g_e = -step_function(ex.speed[0]);
  const double w_0_i = step_function(-in.speed[1]), w_0_e = -step_function(ex.speed[1]);
  const double w_p_i = step_function(-in.speed[2]), w_p_e = -step_function(ex.speed[2]);
  const double w_m_i = step_function(-in.speed[3]), w_m_e = -step_function(ex.speed[3]);
  c_g = w_g_e * ke.v_g - w_g_i * ki.v_g;
  c_pi = 0.5 * (w_p_e * ke.v_plus + w_m_e * ke.v_minus) + w_g_e * ke.g2_v_g -
         0.5 * (w_p_i * ki.v_plus + w_m_i * ki.v_minus) - w_g_i * ki.g2_v_g;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    c_phi[d] = -0.5 * (w_m_e * (ke.v_minus * ex.n_lo[d]) -
                       w_p_e * (ke.v_plus * ex.n_lo[d])) +
               w_0_e * ke.v_zero[d] -
               0.5 * (w_p_i * (ki.v_plus * in.n_lo[d]) -
                      w_m_i * (ki.v_minus * in.n_lo[d])) -
               w_0_i * ki.v_zero[d];
  }
}

// ---------------------------------------------------------------------------
// ScalarWave
// ---------------------------------------------------------------------------
// ScalarWave/TimeDerivative.cpp:35-44 with d_i = J(jhat,i) d_jhat folded in.
// u = (psi, pi, phi_i); d[c][jhat] logical derivatives.
DG_HD void sw_point_rhs(const double (&u)[5], const double (&d)[5][3],
                        const double (&J)[3][3], double gamma2,
                        double (&out)[5]) {
  out[0] = -u[1];
  double di[5][3];
#pragma unroll
  for (int c = 0; c < 5; ++c)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double v = J[0][i] * d[c][0];
      v += J[1][i] * d[c][1];
      v += J[2][i] * d[c][2];
      di[c][i] = v;
    }
  double dtpi = -di[2][0];
  dtpi -= di[3][1];
  dtpi -= di[4][2];
  out[1] = dtpi;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    out[2 + i] = -di[1][i] + gamma2 * (di[0][i] - u[2 + i]);
}

// ScalarWave UpwindPenalty on a static mesh: char speeds (0, +1, -1), so
// w_int = (1, 0, 1), w_ext = (-1, -1, 0) (UpwindPenalty.cpp:36-205).  n_int
// and n_ext are each side's own outward unit normal.
DG_HD void sw_face_correction(const double (&ui)[5], double g2i,
                              const double (&ni)[3], const double (&ue)[5],
                              double g2e, const double (&ne)[3],
                              double (&corr)[5], double ndv_i = 0.0, double ndv_e = 0.0) {
  // packaged data of both sides; ndv = n.v_g of each side's own normal on a moving mesh
  // (UpwindPenalty.cpp:55-67), 0 on a static one
  const double cs0_i = 0.0 - ndv_i, csp_i = 1.0 - ndv_i, csm_i = -1.0 - ndv_i;
  const double cs0_e = 0.0 - ndv_e, csp_e = 1.0 - ndv_e, csm_e = -1.0 - ndv_e;
  double ndphi_i = ni[0] * ui[2];
  ndphi_i += ni[1] * ui[3];
  ndphi_i += ni[2] * ui[4];
  double ndphi_e = ne[0] * ue[2];
  ndphi_e += ne[1] * ue[3];
  ndphi_e += ne[2] * ue[4];
  const double g2psi_i = g2i * ui[0], g2psi_e = g2e * ue[0];
  const double vp_i = csp_i * (ui[1] + ndphi_i - g2psi_i);
  const double vm_i = csm_i * (ui[1] - ndphi_i - g2psi_i);
  const double vp_e = csp_e * (ue[1] + ndphi_e - g2psi_e);
  const double vm_e = csm_e * (ue[1] - ndphi_e - g2psi_e);
  const double w_psi_i = step_function(-cs0_i), w_psi_e = -step_function(cs0_e);
  const double w_p_i = step_function(-csp_i), w_p_e = -step_function(csp_e);
  const double w_m_i = step_function(-csm_i), w_m_e = -step_function(csm_e);
  corr[0] = w_psi_e * (cs0_e * ue[0]) - w_psi_i * (cs0_i * ui[0]);
  corr[1] = 0.5 * (w_p_e * vp_e + w_m_e * vm_e) + w_psi_e * (g2psi_e * cs0_e) -
            0.5 * (w_p_i * vp_i + w_m_i * vm_i) - w_psi_i * (g2psi_i * cs0_i);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double vz_i = cs0_i * (ui[2 + d] - ni[d] * ndphi_i);
    const double vz_e = cs0_e * (ue[2 + d] - ne[d] * ndphi_e);
    corr[2 + d] = 0.5 * (w_p_e * (vp_e * ne[d]) - w_m_e * (vm_e * ne[d])) +
                  w_psi_e * vz_e -
                  0.5 * (w_p_i * (vp_i * ni[d]) - w_m_i * (vm_i * ni[d])) -
                  w_psi_i * vz_i;
  }
}

}  // namespace dg
