/*
 * dg_oracle.c -- CPU restatement of SpECTRE's DG right-hand side for the
 * ScalarWave and GeneralizedHarmonic systems.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  The product path (spectre_b200/) never links, imports
 * or calls anything in oracle/.
 *
 * Every function cites the reference file:line (relative to the reference
 * checkout root, version 2024.09.29) whose algorithm and operation order it
 * restates.  Parity pins (see tests/test_oracle_pins.py):
 *   - gh_rhs_reference_impl vs. the 100 SpEC numbers of
 *     tests/Unit/Evolution/Systems/GeneralizedHarmonic/Test_DuDt.cpp:306-464
 *   - gh_time_derivative vs. gh_rhs_reference_impl on self-consistent input
 *     (same check as Test_DuDt.cpp:466-700)
 *   - package_data / boundary_terms vs. the reference's numpy oracles
 *     tests/Unit/Evolution/Systems/{GeneralizedHarmonic,ScalarWave}/
 *     BoundaryCorrections/UpwindPenalty.py (fixtures in tests/golden/)
 *
 * Layout conventions:
 *   - a Variables block is component-major: comp c of a grid with n points is
 *     v[c*n .. c*n+n); grid index = i + N*(j + N*k)  (xi fastest)
 *   - symmetric spacetime pairs (a<=b) are stored in the order
 *     00 01 02 03 11 12 13 22 23 33 (Tensor/Structure.hpp:162-194)
 *   - GH evolved vars: g_ab (0..9), Pi_ab (10..19), Phi_iab (20 + i + 3*sym)
 *   - SW evolved vars: Psi (0), Pi (1), Phi_i (2..4)
 *   - derivative tensors prepend the derivative index: d_i u_c at 3*c + i
 *   - inverse Jacobian J(ihat, i) = d xi^ihat / d x^i at comp ihat + 3*i
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SYM4(a, b) ((a) <= (b) ? (a) * 4 - (a) * ((a)-1) / 2 + ((b) - (a)) \
                               : (b) * 4 - (b) * ((b)-1) / 2 + ((a) - (b)))
#define SYM3(a, b) ((a) <= (b) ? (a) * 3 - (a) * ((a)-1) / 2 + ((b) - (a)) \
                               : (b) * 3 - (b) * ((b)-1) / 2 + ((a) - (b)))

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* bench.py's CPU arm sets the thread count itself (torchrun exports
 * OMP_NUM_THREADS=1 to its workers) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------
 * logical + inertial partial derivatives
 * NumericalAlgorithms/LinearOperators/PartialDerivatives.tpp:316-363 (three
 * matrix applications over the contiguous Variables block) and :56-110
 * (du = J(0,i) d0 + J(1,i) d1 + J(2,i) d2, first term assigned, rest added).
 * D is row-major D[i*N + j] = D_ij.
 * ---------------------------------------------------------------------- */
void orc_logical_derivs(int N, int C, const double* D, const double* u,
                        double* d0, double* d1, double* d2) {
  const int n = N * N * N;
  for (int c = 0; c < C; ++c) {
    const double* uc = u + (size_t)c * n;
    for (int k = 0; k < N; ++k)
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) {
          double s0 = 0.0, s1 = 0.0, s2 = 0.0;
          for (int m = 0; m < N; ++m) {
            s0 += D[i * N + m] * uc[m + N * (j + N * k)];
            s1 += D[j * N + m] * uc[i + N * (m + N * k)];
            s2 += D[k * N + m] * uc[i + N * (j + N * m)];
          }
          const size_t p = (size_t)c * n + i + N * (j + N * k);
          d0[p] = s0;
          d1[p] = s1;
          d2[p] = s2;
        }
  }
}

void orc_partial_derivatives(int N, int C, const double* D, const double* u,
                             const double* invjac, double* du) {
  const int n = N * N * N;
  double* d0 = (double*)malloc(sizeof(double) * 3 * (size_t)C * n);
  double* d1 = d0 + (size_t)C * n;
  double* d2 = d1 + (size_t)C * n;
  orc_logical_derivs(N, C, D, u, d0, d1, d2);
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < 3; ++i) {
      double* out = du + ((size_t)3 * c + i) * n;
      const double* j0 = invjac + (size_t)(0 + 3 * i) * n;
      const double* j1 = invjac + (size_t)(1 + 3 * i) * n;
      const double* j2 = invjac + (size_t)(2 + 3 * i) * n;
      for (int p = 0; p < n; ++p) {
        double v = j0[p] * d0[(size_t)c * n + p];
        v += j1[p] * d1[(size_t)c * n + p];
        v += j2[p] * d2[(size_t)c * n + p];
        out[p] = v;
      }
    }
  free(d0);
}

/* ------------------------------------------------------------------------
 * ScalarWave::TimeDerivative<3>::apply
 * Evolution/Systems/ScalarWave/TimeDerivative.cpp:14-45
 * ---------------------------------------------------------------------- */
void orc_sw_time_derivative(int n, const double* u, const double* du,
                            const double* gamma2, double* dt) {
  for (int p = 0; p < n; ++p) {
    const double pi = u[1 * n + p];
    dt[0 * n + p] = -pi;
    /* d_phi(d, d): Phi_d is comp 2+d, derivative index d -> 3*(2+d)+d */
    double dtpi = -du[(size_t)(3 * 2 + 0) * n + p];
    dtpi -= du[(size_t)(3 * 3 + 1) * n + p];
    dtpi -= du[(size_t)(3 * 4 + 2) * n + p];
    dt[1 * n + p] = dtpi;
    for (int d = 0; d < 3; ++d) {
      const double d_pi = du[(size_t)(3 * 1 + d) * n + p];
      const double d_psi = du[(size_t)(3 * 0 + d) * n + p];
      dt[(2 + d) * n + p] = -d_pi + gamma2[p] * (d_psi - u[(2 + d) * n + p]);
    }
  }
}

/* ------------------------------------------------------------------------
 * 3x3 symmetric determinant and inverse
 * DataStructures/Tensor/EagerMath/DeterminantAndInverse.hpp:133-160
 * ---------------------------------------------------------------------- */
static void det_and_inverse_sym3(const double t[3][3], double* det,
                                 double inv[3][3]) {
  const double t00 = t[0][0], t01 = t[0][1], t02 = t[0][2];
  const double t11 = t[1][1], t12 = t[1][2], t22 = t[2][2];
  const double a = t11 * t22 - t12 * t12;
  const double b = t12 * t02 - t01 * t22;
  const double c = t01 * t12 - t11 * t02;
  *det = t00 * a + t01 * b + t02 * c;
  const double one_over_det = 1.0 / *det;
  inv[0][0] = (t11 * t22 - t12 * t12) * one_over_det;
  inv[0][1] = inv[1][0] = (t12 * t02 - t22 * t01) * one_over_det;
  inv[0][2] = inv[2][0] = (t01 * t12 - t02 * t11) * one_over_det;
  inv[1][1] = (t22 * t00 - t02 * t02) * one_over_det;
  inv[1][2] = inv[2][1] = (t02 * t01 - t00 * t12) * one_over_det;
  inv[2][2] = (t00 * t11 - t01 * t01) * one_over_det;
}

/* quantities gh::TimeDerivative computes from the metric, exposed for tests */
typedef struct {
  double lapse, shift[3], inv_gamma[3][3], det_gamma, inv_g[4][4];
  double normal_vec[4];
} GhGeom;

/* PointwiseFunctions/GeneralRelativity/{Shift.cpp:27-39, Lapse.cpp:26-34,
 * InverseSpacetimeMetric.cpp:27-48, SpacetimeNormalVector.cpp:28-38} in the
 * order gh::TimeDerivative calls them (TimeDerivative.cpp:87-99,129). */
static void gh_geometry(const double g[4][4], GhGeom* q) {
  double gam[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) gam[i][j] = g[i + 1][j + 1];
  det_and_inverse_sym3(gam, &q->det_gamma, q->inv_gamma);
  for (int i = 0; i < 3; ++i) {
    q->shift[i] = q->inv_gamma[i][0] * g[1][0];
    for (int j = 1; j < 3; ++j) q->shift[i] += q->inv_gamma[i][j] * g[j + 1][0];
  }
  double l = -g[0][0];
  for (int i = 0; i < 3; ++i) l += q->shift[i] * g[i + 1][0];
  q->lapse = sqrt(l);
  const double m1ol2 = -1.0 / (q->lapse * q->lapse);
  q->inv_g[0][0] = m1ol2;
  for (int i = 0; i < 3; ++i)
    q->inv_g[0][i + 1] = q->inv_g[i + 1][0] = -q->shift[i] * m1ol2;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      q->inv_g[i + 1][j + 1] =
          q->inv_gamma[i][j] + q->shift[i] * q->shift[j] * m1ol2;
  q->normal_vec[0] = 1.0 / q->lapse;
  for (int i = 0; i < 3; ++i)
    q->normal_vec[i + 1] = -q->shift[i] * q->normal_vec[0];
}

/* DampedHarmonic parameters (GaugeSourceFunctions/DampedHarmonic.hpp):
 * width sigma_r, amplitudes (L1, L2, S), exponents (L1, L2, S). */
typedef struct {
  double width;
  double amp[3];
  int exp[3];
} DampedHarmonicParams;

static void damped_harmonic_gauge(
    const DampedHarmonicParams* prm, const double x[3], double lapse,
    const double shift[3], double sqrt_det_gamma, const double inv_gamma[3][3],
    const double d4_g[4][4][4], double half_pi_two_normals,
    const double half_phi_two_normals[3], const double g[4][4],
    const double phi[3][4][4], double H[4], double d4H[4][4]);

/* ------------------------------------------------------------------------
 * gh::TimeDerivative<3>::apply at one grid point
 * Evolution/Systems/GeneralizedHarmonic/TimeDerivative.cpp:82-406
 * gauge_mode: 0 = Harmonic (H=0, terms skipped, :255-261,:338-348)
 *             1 = H_a, d_a H_b supplied by the caller (AnalyticChristoffel on
 *                 static data: GaugeSourceFunctions/AnalyticChristoffel.cpp)
 *             2 = DampedHarmonic
 * d4H index convention: d4H[a + 4*b] = d_a H_b (tnsr::ab, first index fastest)
 * ---------------------------------------------------------------------- */
typedef struct {
  int gauge_mode;
  DampedHarmonicParams dh;
} GhGaugeSpec;

static void gh_point(const double g[4][4], const double pi[4][4],
                     const double phi[3][4][4], const double dg[3][4][4],
                     const double dpi[3][4][4], const double dphi[3][3][4][4],
                     double gamma0, double gamma1, double gamma2,
                     const GhGaugeSpec* gauge, const double Hin[4],
                     const double dHin[16], const double x[3],
                     double dt_g[4][4], double dt_pi[4][4],
                     double dt_phi[3][4][4]) {
  GhGeom q;
  gh_geometry(g, &q);
  const double lapse = q.lapse;
  const double* shift = q.shift;

  /* :103-110 provisional dt_g */
  for (int mu = 0; mu < 4; ++mu)
    for (int nu = mu; nu < 4; ++nu) {
      double v = -lapse * pi[mu][nu];
      for (int m = 0; m < 3; ++m) v += shift[m] * phi[m][mu][nu];
      dt_g[mu][nu] = dt_g[nu][mu] = v;
    }
  /* :112-123 da_g: index 0 = provisional dt_g, i+1 = phi_i */
  double da_g[4][4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      da_g[0][a][b] = dt_g[a][b];
      for (int i = 0; i < 3; ++i) da_g[i + 1][a][b] = phi[i][a][b];
    }
  /* :125-128 Christoffel first kind (Christoffel.cpp:16-31), trace */
  double chr1[4][4][4];
  for (int k = 0; k < 4; ++k)
    for (int i = 0; i < 4; ++i)
      for (int j = i; j < 4; ++j) {
        chr1[k][i][j] = chr1[k][j][i] =
            0.5 * (da_g[i][j][k] + da_g[j][i][k] - da_g[k][i][j]);
      }
  double trace_chr[4];
  for (int a = 0; a < 4; ++a) {
    double s = 0.0;
    for (int b = 0; b < 4; ++b)
      for (int c = 0; c < 4; ++c) s += chr1[a][b][c] * q.inv_g[b][c];
    trace_chr[a] = s;
  }
  const double* nv = q.normal_vec;
  const double gamma12 = gamma1 * gamma2;

  /* :134-190 raised quantities */
  double phi_1_up[3][4][4], phi_3_up[3][4][4], pi_2_up[4][4], chr_3_up[4][4][4];
  for (int m = 0; m < 3; ++m)
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = mu; nu < 4; ++nu) {
        double v = q.inv_gamma[m][0] * phi[0][mu][nu];
        for (int n = 1; n < 3; ++n) v += q.inv_gamma[m][n] * phi[n][mu][nu];
        phi_1_up[m][mu][nu] = phi_1_up[m][nu][mu] = v;
      }
  for (int m = 0; m < 3; ++m)
    for (int nu = 0; nu < 4; ++nu)
      for (int al = 0; al < 4; ++al) {
        double v = q.inv_g[al][0] * phi[m][nu][0];
        for (int be = 1; be < 4; ++be) v += q.inv_g[al][be] * phi[m][nu][be];
        phi_3_up[m][nu][al] = v;
      }
  for (int nu = 0; nu < 4; ++nu)
    for (int al = 0; al < 4; ++al) {
      double v = q.inv_g[al][0] * pi[nu][0];
      for (int be = 1; be < 4; ++be) v += q.inv_g[al][be] * pi[nu][be];
      pi_2_up[nu][al] = v;
    }
  for (int mu = 0; mu < 4; ++mu)
    for (int nu = 0; nu < 4; ++nu)
      for (int al = 0; al < 4; ++al) {
        double v = q.inv_g[al][0] * chr1[mu][nu][0];
        for (int be = 1; be < 4; ++be) v += q.inv_g[al][be] * chr1[mu][nu][be];
        chr_3_up[mu][nu][al] = v;
      }
  /* :192-231 normal contractions */
  double pi_one_normal[4];
  for (int mu = 0; mu < 4; ++mu) {
    double v = nv[0] * pi[0][mu];
    for (int nu = 1; nu < 4; ++nu) v += nv[nu] * pi[nu][mu];
    pi_one_normal[mu] = v;
  }
  double half_pi_two_normals = nv[0] * pi_one_normal[0];
  for (int mu = 1; mu < 4; ++mu) half_pi_two_normals += nv[mu] * pi_one_normal[mu];
  half_pi_two_normals *= 0.5;
  double phi_one_normal[3][4], half_phi_two_normals[3];
  for (int n = 0; n < 3; ++n)
    for (int nu = 0; nu < 4; ++nu) {
      double v = nv[0] * phi[n][0][nu];
      for (int mu = 1; mu < 4; ++mu) v += nv[mu] * phi[n][mu][nu];
      phi_one_normal[n][nu] = v;
    }
  for (int n = 0; n < 3; ++n) {
    double v = nv[0] * phi_one_normal[n][0];
    for (int mu = 1; mu < 4; ++mu) v += nv[mu] * phi_one_normal[n][mu];
    half_phi_two_normals[n] = v * 0.5;
  }
  /* :233-240 three-index constraint */
  double c3[3][4][4];
  for (int n = 0; n < 3; ++n)
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = 0; nu < 4; ++nu) c3[n][mu][nu] = dg[n][mu][nu] - phi[n][mu][nu];
  const double gamma1p1 = 1.0 + gamma1;
  /* :245-263 gauge constraint (trace part) and shift . C */
  double gauge_constraint[4], shift_dot_c3[4][4];
  for (int mu = 0; mu < 4; ++mu) {
    gauge_constraint[mu] = trace_chr[mu];
    for (int nu = mu; nu < 4; ++nu) {
      double v = shift[0] * c3[0][mu][nu];
      for (int m = 1; m < 3; ++m) v += shift[m] * c3[m][mu][nu];
      shift_dot_c3[mu][nu] = shift_dot_c3[nu][mu] = v;
    }
  }
  /* :265-286 gauge */
  const int harmonic = (gauge->gauge_mode == 0);
  double H[4] = {0, 0, 0, 0}, dH[4][4];
  double chr2[4][4][4];
  memset(dH, 0, sizeof(dH));
  if (!harmonic) {
    const double sqrt_det = sqrt(q.det_gamma);
    /* raise_or_lower_first_index: Gamma^a_bc = g^{ad} Gamma_dbc */
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b)
        for (int c = b; c < 4; ++c) {
          double v = 0.0;
          for (int d = 0; d < 4; ++d) v += q.inv_g[a][d] * chr1[d][b][c];
          chr2[a][b][c] = chr2[a][c][b] = v;
        }
    if (gauge->gauge_mode == 1) {
      for (int a = 0; a < 4; ++a) H[a] = Hin[a];
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) dH[a][b] = dHin[a + 4 * b];
    } else {
      damped_harmonic_gauge(&gauge->dh, x, lapse, shift, sqrt_det, q.inv_gamma,
                            da_g, half_pi_two_normals, half_phi_two_normals, g,
                            phi, H, dH);
    }
    for (int nu = 0; nu < 4; ++nu) gauge_constraint[nu] += H[nu];
  }
  double normal_dot_gc = nv[0] * gauge_constraint[0];
  for (int mu = 1; mu < 4; ++mu) normal_dot_gc += nv[mu] * gauge_constraint[mu];

  /* :296-306 dt_g */
  for (int mu = 0; mu < 4; ++mu)
    for (int nu = mu; nu < 4; ++nu) {
      dt_g[mu][nu] += gamma1p1 * shift_dot_c3[mu][nu];
      dt_g[nu][mu] = dt_g[mu][nu];
    }
  /* :308-340 dt_pi, n_a pieces (normal_dot_gc rescaled by gamma0) */
  normal_dot_gc *= gamma0;
  dt_pi[0][0] = -gamma0 * lapse;
  for (int i = 1; i < 4; ++i)
    dt_pi[0][i] = dt_pi[0][0] * gauge_constraint[i] - normal_dot_gc * g[0][i];
  dt_pi[0][0] = 2.0 * dt_pi[0][0] * gauge_constraint[0] - normal_dot_gc * g[0][0];
  for (int mu = 1; mu < 4; ++mu)
    for (int nu = mu; nu < 4; ++nu) dt_pi[mu][nu] = -normal_dot_gc * g[mu][nu];
  /* :342-392 */
  for (int mu = 0; mu < 4; ++mu)
    for (int nu = mu; nu < 4; ++nu) {
      double v = dt_pi[mu][nu];
      v -= half_pi_two_normals * pi[mu][nu];
      if (!harmonic) v -= dH[mu][nu] + dH[nu][mu];
      for (int de = 0; de < 4; ++de) {
        v -= 2 * pi[mu][de] * pi_2_up[nu][de];
        if (!harmonic) v += 2 * chr2[de][mu][nu] * H[de];
        for (int n = 0; n < 3; ++n)
          v += 2 * phi_1_up[n][mu][de] * phi_3_up[n][nu][de];
        for (int al = 0; al < 4; ++al)
          v -= 2. * chr_3_up[mu][al][de] * chr_3_up[nu][de][al];
      }
      for (int m = 0; m < 3; ++m) {
        v -= pi_one_normal[m + 1] * phi_1_up[m][mu][nu];
        for (int n = 0; n < 3; ++n) v -= q.inv_gamma[m][n] * dphi[m][n][mu][nu];
      }
      v *= lapse;
      v += gamma12 * shift_dot_c3[mu][nu];
      for (int m = 0; m < 3; ++m) v += shift[m] * dpi[m][mu][nu];
      dt_pi[mu][nu] = dt_pi[nu][mu] = v;
    }
  /* :394-412 dt_phi */
  for (int i = 0; i < 3; ++i)
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = mu; nu < 4; ++nu) {
        double v = pi[mu][nu] * half_phi_two_normals[i] - dpi[i][mu][nu] +
                   gamma2 * c3[i][mu][nu];
        for (int n = 0; n < 3; ++n)
          v += phi_one_normal[i][n + 1] * phi_1_up[n][mu][nu];
        v *= lapse;
        for (int m = 0; m < 3; ++m) v += shift[m] * dphi[m][i][mu][nu];
        dt_phi[i][mu][nu] = dt_phi[i][nu][mu] = v;
      }
}

/* gather/scatter helpers between SoA Variables layout and per-point arrays */
static void gh_gather(int n, int p, const double* u, double g[4][4],
                      double pi[4][4], double phi[3][4][4]) {
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      const int s = SYM4(a, b);
      g[a][b] = u[(size_t)s * n + p];
      pi[a][b] = u[(size_t)(10 + s) * n + p];
      for (int i = 0; i < 3; ++i) phi[i][a][b] = u[(size_t)(20 + i + 3 * s) * n + p];
    }
}

static void gh_gather_derivs(int n, int p, const double* du, double dg[3][4][4],
                             double dpi[3][4][4], double dphi[3][3][4][4]) {
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      const int s = SYM4(a, b);
      for (int i = 0; i < 3; ++i) {
        dg[i][a][b] = du[(size_t)(3 * s + i) * n + p];
        dpi[i][a][b] = du[(size_t)(3 * (10 + s) + i) * n + p];
        for (int j = 0; j < 3; ++j)
          dphi[i][j][a][b] = du[(size_t)(3 * (20 + j + 3 * s) + i) * n + p];
      }
    }
}

/* gauge_params: [mode, width, A_L1, A_L2, A_S, e_L1, e_L2, e_S] */
static void gauge_from_params(const double* gp, GhGaugeSpec* gs) {
  gs->gauge_mode = (int)gp[0];
  gs->dh.width = gp[1];
  for (int i = 0; i < 3; ++i) {
    gs->dh.amp[i] = gp[2 + i];
    gs->dh.exp[i] = (int)gp[5 + i];
  }
}

void orc_gh_time_derivative(int n, const double* u, const double* du,
                            const double* gamma0, const double* gamma1,
                            const double* gamma2, const double* gauge_params,
                            const double* H, const double* dH,
                            const double* coords, double* dt) {
  GhGaugeSpec gs;
  gauge_from_params(gauge_params, &gs);
  for (int p = 0; p < n; ++p) {
    double g[4][4], pi[4][4], phi[3][4][4], dg[3][4][4], dpi[3][4][4],
        dphi[3][3][4][4], dtg[4][4], dtpi[4][4], dtphi[3][4][4];
    double Hp[4] = {0, 0, 0, 0}, dHp[16] = {0};
    double x[3] = {0, 0, 0};
    gh_gather(n, p, u, g, pi, phi);
    gh_gather_derivs(n, p, du, dg, dpi, dphi);
    if (gs.gauge_mode == 1) {
      for (int a = 0; a < 4; ++a) Hp[a] = H[(size_t)a * n + p];
      for (int a = 0; a < 16; ++a) dHp[a] = dH[(size_t)a * n + p];
    }
    if (coords)
      for (int i = 0; i < 3; ++i) x[i] = coords[(size_t)i * n + p];
    gh_point(g, pi, phi, dg, dpi, dphi, gamma0[p], gamma1[p], gamma2[p], &gs,
             Hp, dHp, x, dtg, dtpi, dtphi);
    for (int a = 0; a < 4; ++a)
      for (int b = a; b < 4; ++b) {
        const int s = SYM4(a, b);
        dt[(size_t)s * n + p] = dtg[a][b];
        dt[(size_t)(10 + s) * n + p] = dtpi[a][b];
        for (int i = 0; i < 3; ++i)
          dt[(size_t)(20 + i + 3 * s) * n + p] = dtphi[i][a][b];
      }
  }
}

/* geometry exposed for tests: out = [lapse, shift(3), inv_gamma sym(6),
 * inv_g sym(10), det_gamma] per point, component-major */
void orc_gh_geometry(int n, const double* u, double* out) {
  for (int p = 0; p < n; ++p) {
    double g[4][4], pi[4][4], phi[3][4][4];
    GhGeom q;
    gh_gather(n, p, u, g, pi, phi);
    gh_geometry(g, &q);
    int c = 0;
    out[(size_t)(c++) * n + p] = q.lapse;
    for (int i = 0; i < 3; ++i) out[(size_t)(c++) * n + p] = q.shift[i];
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) out[(size_t)(c++) * n + p] = q.inv_gamma[i][j];
    for (int a = 0; a < 4; ++a)
      for (int b = a; b < 4; ++b) out[(size_t)(c++) * n + p] = q.inv_g[a][b];
    out[(size_t)(c++) * n + p] = q.det_gamma;
  }
}

/* ------------------------------------------------------------------------
 * gh_rhs_reference_impl of the reference's own test
 * tests/Unit/Evolution/Systems/GeneralizedHarmonic/Test_DuDt.cpp:51-290
 * (takes every geometric quantity as an independent input so that the SpEC
 * numbers at :353-464 can be reproduced).  All tensors at ONE point, full
 * (unsymmetrised-index) arrays.
 * ---------------------------------------------------------------------- */
void orc_gh_rhs_reference_impl(
    const double* g_, const double* pi_, const double* phi_, const double* dg_,
    const double* dpi_, const double* dphi_, double gamma0, double gamma1,
    double gamma2, const double* H, const double* dH_, double lapse,
    const double* shift, const double* inv_gamma_, const double* inv_g_,
    const double* trace_chr, const double* chr1_, const double* chr2_,
    const double* nvec, const double* nform, double* dt_g_, double* dt_pi_,
    double* dt_phi_) {
  /* inputs are dense C arrays: g[4][4], phi[3][4][4], dg[3][4][4],
   * dphi[3][3][4][4] (deriv index first), dH[4][4] (dH[a][b] = d_a H_b),
   * inv_gamma[3][3], inv_g[4][4], chr1[4][4][4], chr2[4][4][4] */
  const double(*g)[4] = (const double(*)[4])g_;
  const double(*pi)[4] = (const double(*)[4])pi_;
  const double(*phi)[4][4] = (const double(*)[4][4])phi_;
  const double(*dg)[4][4] = (const double(*)[4][4])dg_;
  const double(*dpi)[4][4] = (const double(*)[4][4])dpi_;
  const double(*dphi)[3][4][4] = (const double(*)[3][4][4])dphi_;
  const double(*dH)[4] = (const double(*)[4])dH_;
  const double(*inv_gamma)[3] = (const double(*)[3])inv_gamma_;
  const double(*inv_g)[4] = (const double(*)[4])inv_g_;
  const double(*chr1)[4][4] = (const double(*)[4][4])chr1_;
  const double(*chr2)[4][4] = (const double(*)[4][4])chr2_;
  double(*dt_g)[4] = (double(*)[4])dt_g_;
  double(*dt_pi)[4] = (double(*)[4])dt_pi_;
  double(*dt_phi)[4][4] = (double(*)[4][4])dt_phi_;

  const double gamma12 = gamma1 * gamma2;
  double phi_1_up[3][4][4] = {{{0}}}, phi_3_up[3][4][4] = {{{0}}},
         pi_2_up[4][4] = {{0}}, chr_3_up[4][4][4] = {{{0}}};
  for (int m = 0; m < 3; ++m)
    for (int mu = 0; mu < 4; ++mu)
      for (int n = 0; n < 3; ++n)
        for (int nu = mu; nu < 4; ++nu)
          phi_1_up[m][mu][nu] += inv_gamma[m][n] * phi[n][mu][nu];
  for (int m = 0; m < 3; ++m)
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = mu; nu < 4; ++nu) phi_1_up[m][nu][mu] = phi_1_up[m][mu][nu];
  for (int m = 0; m < 3; ++m)
    for (int nu = 0; nu < 4; ++nu)
      for (int al = 0; al < 4; ++al)
        for (int be = 0; be < 4; ++be)
          phi_3_up[m][nu][al] += inv_g[al][be] * phi[m][nu][be];
  for (int nu = 0; nu < 4; ++nu)
    for (int al = 0; al < 4; ++al)
      for (int be = 0; be < 4; ++be) pi_2_up[nu][al] += inv_g[al][be] * pi[nu][be];
  for (int mu = 0; mu < 4; ++mu)
    for (int nu = 0; nu < 4; ++nu)
      for (int al = 0; al < 4; ++al)
        for (int be = 0; be < 4; ++be)
          chr_3_up[mu][nu][al] += inv_g[al][be] * chr1[mu][nu][be];
  double pi_dot_n[4] = {0}, pi_nn = 0.0, phi_dot_n[3][4] = {{0}}, phi_nn[3] = {0};
  for (int nu = 0; nu < 4; ++nu)
    for (int mu = 0; mu < 4; ++mu) pi_dot_n[mu] += nvec[nu] * pi[nu][mu];
  for (int mu = 0; mu < 4; ++mu) pi_nn += nvec[mu] * pi_dot_n[mu];
  for (int n = 0; n < 3; ++n)
    for (int nu = 0; nu < 4; ++nu)
      for (int mu = 0; mu < 4; ++mu) phi_dot_n[n][nu] += nvec[mu] * phi[n][mu][nu];
  for (int n = 0; n < 3; ++n)
    for (int mu = 0; mu < 4; ++mu) phi_nn[n] += nvec[mu] * phi_dot_n[n][mu];
  double c3[3][4][4], c1[4], n_dot_c1 = 0.0, sc3[4][4] = {{0}};
  for (int n = 0; n < 3; ++n)
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = 0; nu < 4; ++nu) c3[n][mu][nu] = dg[n][mu][nu] - phi[n][mu][nu];
  for (int nu = 0; nu < 4; ++nu) c1[nu] = H[nu] + trace_chr[nu];
  for (int mu = 0; mu < 4; ++mu) n_dot_c1 += nvec[mu] * c1[mu];
  const double gamma1p1 = 1.0 + gamma1;
  for (int m = 0; m < 3; ++m)
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = mu; nu < 4; ++nu) sc3[mu][nu] += shift[m] * c3[m][mu][nu];
  for (int mu = 0; mu < 4; ++mu)
    for (int nu = mu; nu < 4; ++nu) {
      double v = -lapse * pi[mu][nu];
      v += gamma1p1 * sc3[mu][nu];
      for (int m = 0; m < 3; ++m) v += shift[m] * phi[m][mu][nu];
      dt_g[mu][nu] = dt_g[nu][mu] = v;
    }
  for (int mu = 0; mu < 4; ++mu)
    for (int nu = mu; nu < 4; ++nu) {
      double v = -dH[mu][nu] - dH[nu][mu] - 0.5 * pi_nn * pi[mu][nu] +
                 gamma0 * (nform[mu] * c1[nu] + nform[nu] * c1[mu]) -
                 gamma0 * g[mu][nu] * n_dot_c1;
      for (int de = 0; de < 4; ++de) {
        v += 2 * chr2[de][mu][nu] * H[de] - 2 * pi[mu][de] * pi_2_up[nu][de];
        for (int n = 0; n < 3; ++n)
          v += 2 * phi_1_up[n][mu][de] * phi_3_up[n][nu][de];
        for (int al = 0; al < 4; ++al)
          v -= 2. * chr_3_up[mu][al][de] * chr_3_up[nu][de][al];
      }
      for (int m = 0; m < 3; ++m) {
        v -= pi_dot_n[m + 1] * phi_1_up[m][mu][nu];
        for (int n = 0; n < 3; ++n) v -= inv_gamma[m][n] * dphi[m][n][mu][nu];
      }
      v *= lapse;
      v += gamma12 * sc3[mu][nu];
      for (int m = 0; m < 3; ++m) v += shift[m] * dpi[m][mu][nu];
      dt_pi[mu][nu] = dt_pi[nu][mu] = v;
    }
  for (int i = 0; i < 3; ++i)
    for (int mu = 0; mu < 4; ++mu)
      for (int nu = mu; nu < 4; ++nu) {
        double v = 0.5 * pi[mu][nu] * phi_nn[i] - dpi[i][mu][nu] +
                   gamma2 * c3[i][mu][nu];
        for (int n = 0; n < 3; ++n) v += phi_dot_n[i][n + 1] * phi_1_up[n][mu][nu];
        v *= lapse;
        for (int m = 0; m < 3; ++m) v += shift[m] * dphi[m][i][mu][nu];
        dt_phi[i][mu][nu] = dt_phi[i][nu][mu] = v;
      }
}

/* ------------------------------------------------------------------------
 * Face normals
 * Evolution/DiscontinuousGalerkin/Actions/InternalMortarDataImpl.hpp:180-221
 * (unnormalised covector = +-row `dim` of the inverse Jacobian on the face)
 * and NormalCovectorAndMagnitude.hpp:47-92 (curved: n^i = gamma^{ij} n_j,
 * |n| = sqrt(n^i n_i); flat: Euclidean magnitude).
 * ---------------------------------------------------------------------- */
static void face_normal(const double unnorm[3], int curved,
                        const double inv_gamma[3][3], double n_lo[3],
                        double n_up[3], double* mag) {
  if (curved) {
    for (int i = 0; i < 3; ++i) {
      n_up[i] = inv_gamma[i][0] * unnorm[0];
      for (int j = 1; j < 3; ++j) n_up[i] += inv_gamma[i][j] * unnorm[j];
    }
    double m = n_up[0] * unnorm[0];
    for (int i = 1; i < 3; ++i) m += n_up[i] * unnorm[i];
    *mag = sqrt(m);
    const double inv = 1.0 / *mag;
    for (int i = 0; i < 3; ++i) {
      n_lo[i] = unnorm[i] * inv;
      n_up[i] *= inv;
    }
  } else {
    double m = unnorm[0] * unnorm[0];
    for (int i = 1; i < 3; ++i) m += unnorm[i] * unnorm[i];
    *mag = sqrt(m);
    const double inv = 1.0 / *mag;
    for (int i = 0; i < 3; ++i) {
      n_lo[i] = unnorm[i] * inv;
      n_up[i] = n_lo[i];
    }
  }
}

/* ------------------------------------------------------------------------
 * ScalarWave UpwindPenalty, one face point.
 * packaged (16): v_psi, v_zero(3), v_plus, v_minus, n_times_v_plus(3),
 * n_times_v_minus(3), gamma2_v_psi, char_speeds(3)
 * Evolution/Systems/ScalarWave/BoundaryCorrections/UpwindPenalty.cpp:36-108
 * ---------------------------------------------------------------------- */
static void sw_package_point_moving(const double u[5], double gamma2,
                                    const double n[3], double ndotv, double pk[16]) {
  const double psi = u[0], pi = u[1];
  const double* phi = u + 2;
  double* cs = pk + 13;
  /* UpwindPenalty.cpp:55-67: the mesh moves with n.v_g along the normal */
  cs[0] = 0.0 - ndotv;
  cs[1] = 1.0 - ndotv;
  cs[2] = -1.0 - ndotv;
  double g2psi = gamma2 * psi;
  double ndphi = n[0] * phi[0];
  ndphi += n[1] * phi[1];
  ndphi += n[2] * phi[2];
  for (int i = 0; i < 3; ++i) pk[1 + i] = cs[0] * (phi[i] - n[i] * ndphi);
  pk[4] = cs[1] * (pi + ndphi - g2psi);
  pk[5] = cs[2] * (pi - ndphi - g2psi);
  for (int d = 0; d < 3; ++d) {
    pk[6 + d] = pk[4] * n[d];
    pk[9 + d] = pk[5] * n[d];
  }
  pk[0] = cs[0] * psi;
  pk[12] = g2psi * cs[0];
}
static void sw_package_point(const double u[5], double gamma2,
                             const double n[3], double pk[16]) {
  sw_package_point_moving(u, gamma2, n, 0.0, pk);
}

static double step_function(double x) { return x < 0.0 ? 0.0 : 1.0; }

/* ScalarWave/BoundaryCorrections/UpwindPenalty.cpp:111-205 */
static void sw_boundary_terms_point(const double in[16], const double ex[16],
                                    double corr[5]) {
  const double* csi = in + 13;
  const double* cse = ex + 13;
  const double w_psi_i = step_function(-csi[0]), w_psi_e = -step_function(cse[0]);
  const double w_0_i = step_function(-csi[0]), w_0_e = -step_function(cse[0]);
  const double w_p_i = step_function(-csi[1]), w_p_e = -step_function(cse[1]);
  const double w_m_i = step_function(-csi[2]), w_m_e = -step_function(cse[2]);
  corr[0] = w_psi_e * ex[0] - w_psi_i * in[0];
  corr[1] = 0.5 * (w_p_e * ex[4] + w_m_e * ex[5]) + w_psi_e * ex[12] -
            0.5 * (w_p_i * in[4] + w_m_i * in[5]) - w_psi_i * in[12];
  for (int d = 0; d < 3; ++d)
    corr[2 + d] = 0.5 * (w_p_e * ex[6 + d] - w_m_e * ex[9 + d]) + w_0_e * ex[1 + d] -
                  0.5 * (w_p_i * in[6 + d] - w_m_i * in[9 + d]) - w_0_i * in[1 + d];
}

void orc_sw_package_data(int f, const double* u, const double* gamma2,
                         const double* normal, double* packaged) {
  for (int p = 0; p < f; ++p) {
    double up[5], n[3], pk[16];
    for (int c = 0; c < 5; ++c) up[c] = u[(size_t)c * f + p];
    for (int i = 0; i < 3; ++i) n[i] = normal[(size_t)i * f + p];
    sw_package_point(up, gamma2[p], n, pk);
    for (int c = 0; c < 16; ++c) packaged[(size_t)c * f + p] = pk[c];
  }
}

void orc_sw_boundary_terms(int f, const double* pk_int, const double* pk_ext,
                           double* corr) {
  for (int p = 0; p < f; ++p) {
    double in[16], ex[16], c[5];
    for (int k = 0; k < 16; ++k) {
      in[k] = pk_int[(size_t)k * f + p];
      ex[k] = pk_ext[(size_t)k * f + p];
    }
    sw_boundary_terms_point(in, ex, c);
    for (int k = 0; k < 5; ++k) corr[(size_t)k * f + p] = c[k];
  }
}

/* ------------------------------------------------------------------------
 * GH UpwindPenalty, one face point.  packaged (134), in the reference's
 * dg_package_field_tags order (UpwindPenalty.hpp:215-226):
 *   v_g (10) | v_zero (30, i + 3*sym) | v_plus (10) | v_minus (10) |
 *   n_times_v_plus (30) | n_times_v_minus (30) | gamma2_v_g (10) | speeds (4)
 * Evolution/Systems/GeneralizedHarmonic/BoundaryCorrections/
 *   UpwindPenalty.cpp:36-158
 * ---------------------------------------------------------------------- */
static void gh_package_point_moving(const double u[50], double gamma1, double gamma2,
                                    double lapse, const double shift[3],
                                    const double n_lo[3], const double n_up[3],
                                    double ndotv, double pk[134]) {
  double* v_g = pk;
  double* v_zero = pk + 10;
  double* v_plus = pk + 40;
  double* v_minus = pk + 50;
  double* nvp = pk + 60;
  double* nvm = pk + 90;
  double* g2vg = pk + 120;
  double* cs = pk + 130;
  double sdn = shift[0] * n_lo[0];
  sdn += shift[1] * n_lo[1];
  sdn += shift[2] * n_lo[2];
  sdn *= -1.0;
  cs[1] = sdn;
  cs[0] = (1.0 + gamma1) * sdn;
  cs[2] = lapse + sdn;
  cs[3] = -lapse + sdn;
  /* GH UpwindPenalty.cpp:85-91: speeds without the mesh movement, then the mesh movement */
  cs[0] -= ndotv * (1.0 + gamma1);
  cs[1] -= ndotv;
  cs[2] -= ndotv;
  cs[3] -= ndotv;
  for (int s = 0; s < 10; ++s) g2vg[s] = gamma2 * u[s];
  for (int s = 0; s < 10; ++s) {
    double ndphi = n_up[0] * u[20 + 0 + 3 * s];
    for (int i = 1; i < 3; ++i) ndphi += n_up[i] * u[20 + i + 3 * s];
    v_plus[s] = cs[2] * (u[10 + s] + ndphi - g2vg[s]);
    v_minus[s] = cs[3] * (u[10 + s] - ndphi - g2vg[s]);
    for (int i = 0; i < 3; ++i)
      v_zero[i + 3 * s] = cs[1] * (u[20 + i + 3 * s] - n_lo[i] * ndphi);
  }
  for (int s = 0; s < 10; ++s) {
    for (int d = 0; d < 3; ++d) {
      nvp[d + 3 * s] = v_plus[s] * n_lo[d];
      nvm[d + 3 * s] = v_minus[s] * n_lo[d];
    }
    v_g[s] = cs[0] * u[s];
    g2vg[s] *= cs[0];
  }
}

static void gh_package_point(const double u[50], double gamma1, double gamma2,
                             double lapse, const double shift[3],
                             const double n_lo[3], const double n_up[3],
                             double pk[134]) {
  gh_package_point_moving(u, gamma1, gamma2, lapse, shift, n_lo, n_up, 0.0, pk);
}

/* GeneralizedHarmonic/BoundaryCorrections/UpwindPenalty.cpp:161-275 */
static void gh_boundary_terms_point(const double in[134], const double ex[134],
                                    double corr[50]) {
  const double* csi = in + 130;
  const double* cse = ex + 130;
  const double w_g_i = step_function(-csi[0]), w_g_e = -step_function(cse[0]);
  const double w_0_i = step_function(-csi[1]), w_0_e = -step_function(cse[1]);
  const double w_p_i = step_function(-csi[2]), w_p_e = -step_function(cse[2]);
  const double w_m_i = step_function(-csi[3]), w_m_e = -step_function(cse[3]);
  for (int s = 0; s < 10; ++s) {
    corr[s] = w_g_e * ex[s] - w_g_i * in[s];
    corr[10 + s] = 0.5 * (w_p_e * ex[40 + s] + w_m_e * ex[50 + s]) +
                   w_g_e * ex[120 + s] -
                   0.5 * (w_p_i * in[40 + s] + w_m_i * in[50 + s]) -
                   w_g_i * in[120 + s];
    for (int d = 0; d < 3; ++d) {
      const int k = d + 3 * s;
      corr[20 + k] = -0.5 * (w_m_e * ex[90 + k] - w_p_e * ex[60 + k]) +
                     w_0_e * ex[10 + k] -
                     0.5 * (w_p_i * in[60 + k] - w_m_i * in[90 + k]) -
                     w_0_i * in[10 + k];
    }
  }
}

void orc_gh_package_data(int f, const double* u, const double* gamma1,
                         const double* gamma2, const double* lapse,
                         const double* shift, const double* n_lo,
                         const double* n_up, double* packaged) {
  for (int p = 0; p < f; ++p) {
    double up[50], sh[3], nl[3], nu[3], pk[134];
    for (int c = 0; c < 50; ++c) up[c] = u[(size_t)c * f + p];
    for (int i = 0; i < 3; ++i) {
      sh[i] = shift[(size_t)i * f + p];
      nl[i] = n_lo[(size_t)i * f + p];
      nu[i] = n_up[(size_t)i * f + p];
    }
    gh_package_point(up, gamma1[p], gamma2[p], lapse[p], sh, nl, nu, pk);
    for (int c = 0; c < 134; ++c) packaged[(size_t)c * f + p] = pk[c];
  }
}

void orc_gh_package_data_moving(int f, const double* u, const double* gamma1,
                                const double* gamma2, const double* lapse,
                                const double* shift, const double* n_lo,
                                const double* n_up, const double* ndotv, double* packaged) {
  for (int p = 0; p < f; ++p) {
    double up[50], sh[3], nl[3], nu[3], pk[134];
    for (int c = 0; c < 50; ++c) up[c] = u[(size_t)c * f + p];
    for (int i = 0; i < 3; ++i) {
      sh[i] = shift[(size_t)i * f + p];
      nl[i] = n_lo[(size_t)i * f + p];
      nu[i] = n_up[(size_t)i * f + p];
    }
    gh_package_point_moving(up, gamma1[p], gamma2[p], lapse[p], sh, nl, nu, ndotv[p], pk);
    for (int c = 0; c < 134; ++c) packaged[(size_t)c * f + p] = pk[c];
  }
}

void orc_sw_package_data_moving(int f, const double* u, const double* gamma2,
                                const double* normal, const double* ndotv, double* packaged) {
  for (int p = 0; p < f; ++p) {
    double up[5], n[3], pk[16];
    for (int c = 0; c < 5; ++c) up[c] = u[(size_t)c * f + p];
    for (int i = 0; i < 3; ++i) n[i] = normal[(size_t)i * f + p];
    sw_package_point_moving(up, gamma2[p], n, ndotv[p], pk);
    for (int c = 0; c < 16; ++c) packaged[(size_t)c * f + p] = pk[c];
  }
}

/* Mesh velocity of the next orc_dg_rhs* call ([nelem][3][n], inertial components; NULL =
 * static mesh): VolumeTermsImpl.tpp:155-235 (dt u += v_g^i d_i u for systems without fluxes),
 * GeneralizedHarmonic/TimeDerivative.cpp:237-300,372-378 (gamma1 v_g.C3 in dt g,
 * gamma1 gamma2 v_g.C3 in dt Pi), normal_dot_mesh_velocity in dg_package_data. */
static const double* g_mesh_velocity = NULL;
void orc_set_mesh_velocity(const double* v) { g_mesh_velocity = v; }

void orc_gh_boundary_terms(int f, const double* pk_int, const double* pk_ext,
                           double* corr) {
  for (int p = 0; p < f; ++p) {
    double in[134], ex[134], c[50];
    for (int k = 0; k < 134; ++k) {
      in[k] = pk_int[(size_t)k * f + p];
      ex[k] = pk_ext[(size_t)k * f + p];
    }
    gh_boundary_terms_point(in, ex, c);
    for (int k = 0; k < 50; ++k) corr[(size_t)k * f + p] = c[k];
  }
}

/* ------------------------------------------------------------------------
 * Whole-domain DG right-hand side, GTS, conforming aligned mortars (Brick).
 * Order of operations per element follows
 *   ComputeTimeDerivative.hpp:383-650 -> volume_terms (VolumeTermsImpl.tpp:
 *   72-295) -> internal_mortar_data_impl (InternalMortarDataImpl.hpp:45-320:
 *   project to face, normal, dg_package_data) and then
 *   ApplyBoundaryCorrections.hpp:797-1048 (dg_boundary_terms, lift_flux with
 *   -0.5*N*(N-1)*|n|, add_slice_to_data).
 * system: 0 = ScalarWave (C=5, PK=16), 1 = GH (C=50, PK=134)
 * nbr[e*6 + d]: neighbour element index for direction d (0:-xi 1:+xi 2:-eta
 *   3:+eta 4:-zeta 5:+zeta), aligned orientation; -1 = external boundary
 *   (no correction applied: only periodic domains are covered here).
 * static_fields: SW: gamma2 (1 comp); GH: gamma0,gamma1,gamma2 (3 comps)
 *   then, if gauge mode 1, H (4) and dH (16), per element component-major.
 * The face contributions are added in direction order 0..5 (the reference's
 * order is that of a hash map, i.e. unspecified: SURVEY Appendix A.11).
 * ---------------------------------------------------------------------- */
static int face_index(int N, int d, int a, int b) {
  const int dim = d / 2, side = d % 2;
  const int fixed = side ? N - 1 : 0;
  switch (dim) {
    case 0: return fixed + N * (a + N * b);
    case 1: return a + N * (fixed + N * b);
    default: return a + N * (b + N * fixed);
  }
}

/*
 * External boundaries with a ghost boundary condition (DirichletAnalytic):
 * nbr <= -2 selects slot -(nbr+2) of ext_u [nslots][C][f], the exterior
 * evolved variables returned by BoundaryCondition::dg_ghost
 * (GeneralizedHarmonic/BoundaryConditions/DirichletAnalytic.cpp:58-117).
 * Following BoundaryConditionsImpl.hpp:427-560: gamma1/gamma2 are copied from
 * the interior, lapse/shift/inverse spatial metric come from the exterior
 * metric, and the exterior normal is minus the interior UNIT normal covector
 * re-normalised with the exterior inverse spatial metric.
 */
void orc_dg_rhs_oriented(int system, int N, int nelem, const double* D, const double* u,
                         const double* invjac, const double* static_fields,
                         const double* coords, const int* nbr, const int* nbr_face,
                         const double* gauge_params, const double* ext_u, double* dt_u);
void orc_dg_rhs_mortars(int system, int N, int nelem, const double* D, const double* u,
                        const double* invjac, const double* static_fields,
                        const double* coords, const int* nbr, const int* nbr_face,
                        const double* gauge_params, const double* ext_u, int n_mortars,
                        const int* mortars, const double* P, const double* R,
                        double* dt_u);

void orc_dg_rhs(int system, int N, int nelem, const double* D, const double* u,
                const double* invjac, const double* static_fields,
                const double* coords, const int* nbr,
                const double* gauge_params, double* dt_u) {
  orc_dg_rhs_oriented(system, N, nelem, D, u, invjac, static_fields, coords, nbr, NULL,
                      gauge_params, NULL, dt_u);
}

void orc_dg_rhs_bc(int system, int N, int nelem, const double* D, const double* u,
                   const double* invjac, const double* static_fields,
                   const double* coords, const int* nbr,
                   const double* gauge_params, const double* ext_u, double* dt_u) {
  orc_dg_rhs_oriented(system, N, nelem, D, u, invjac, static_fields, coords, nbr, NULL,
                      gauge_params, ext_u, dt_u);
}

/*
 * Non-aligned neighbours: the packaged data a neighbour sends are re-ordered
 * into the receiver's frame before dg_boundary_terms is called
 * (orient_variables_on_slice, Domain/Structure/OrientationMapHelpers.cpp:25-120;
 * call site ComputeTimeDerivative.hpp:712-721).  All packaged fields are
 * scalars or INERTIAL tensor components, so only the face-point index changes.
 * nbr_face[e*6+d] = nd | (perm << 3): nd = the neighbour's direction touching
 * this face; perm bit 0 = the two face coordinates are swapped, bit 1 / bit 2 =
 * the neighbour's first / second face coordinate runs backwards.  NULL = aligned
 * (nd = d ^ 1, perm = 0).
 */
void orc_dg_rhs_oriented(int system, int N, int nelem, const double* D, const double* u,
                         const double* invjac, const double* static_fields,
                         const double* coords, const int* nbr, const int* nbr_face,
                         const double* gauge_params, const double* ext_u, double* dt_u) {
  orc_dg_rhs_mortars(system, N, nelem, D, u, invjac, static_fields, coords, nbr, nbr_face,
                     gauge_params, ext_u, 0, NULL, NULL, NULL, dt_u);
}

/*
 * Non-conforming (h-refined, 2:1) mortars.  A face whose neighbour table entry
 * is ORC_HANGING is skipped by the conforming loop; its corrections come from
 * the mortar table: row m = {coarse element, its direction, fine element, its
 * direction, size_a, size_b} with the MortarSize of the fine face inside the
 * coarse face per face dimension (0 Full, 1 LowerHalf, 2 UpperHalf;
 * dg::mortar_size, MortarHelpers.cpp:51-77; blocks aligned).  The mortar mesh
 * is the fine face.  Following InternalMortarDataImpl.hpp:230-320 and
 * ApplyBoundaryCorrections.hpp:797-1045: each side packages on its own FACE; the
 * coarse side's packaged data are interpolated to the mortar
 * (project_to_mortar, projection_matrix_parent_to_child: P[size] row-major
 * [child point][parent point]); dg_boundary_terms is evaluated on the mortar for
 * both elements; the coarse element's correction is L2-projected back to its
 * face (project_from_mortar, projection_matrix_child_to_parent: R[size]
 * row-major [parent point][child point]); each side lifts with the normal
 * magnitude on its own face (LiftFlux.hpp:57-61) and adds the slice.
 */
#define ORC_HANGING (-2147483647 - 1)
/* external face with a TimeDerivative-type boundary condition (Bjorhus): no
 * boundary correction here, the caller adds the condition's dt corrections */
#define ORC_BJORHUS (-2147483647)
#define ORC_BJORHUS_PHYSICAL (-2147483646)

void orc_dg_rhs_mortars(int system, int N, int nelem, const double* D, const double* u,
                        const double* invjac, const double* static_fields,
                        const double* coords, const int* nbr, const int* nbr_face,
                        const double* gauge_params, const double* ext_u, int n_mortars,
                        const int* mortars, const double* P, const double* R,
                        double* dt_u) {
  const int n = N * N * N, f = N * N;
  const int C = system == 0 ? 5 : 50;
  const int PK = system == 0 ? 16 : 134;
  GhGaugeSpec gs;
  gs.gauge_mode = 0;
  if (system == 1) gauge_from_params(gauge_params, &gs);
  const int nstatic = system == 0 ? 1 : (gs.gauge_mode == 1 ? 23 : 3);
  /* packaged data for every face of every element */
  double* pk_all = (double*)malloc(sizeof(double) * (size_t)nelem * 6 * PK * f);
  double* mag_all = (double*)malloc(sizeof(double) * (size_t)nelem * 6 * f);
#pragma omp parallel
  {
    double* du = (double*)malloc(sizeof(double) * 3 * (size_t)C * n);
#pragma omp for schedule(static)
    for (int e = 0; e < nelem; ++e) {
      const double* ue = u + (size_t)e * C * n;
      const double* je = invjac + (size_t)e * 9 * n;
      const double* se = static_fields + (size_t)e * nstatic * n;
      double* dte = dt_u + (size_t)e * C * n;
      orc_partial_derivatives(N, C, D, ue, je, du);
      if (system == 0) {
        orc_sw_time_derivative(n, ue, du, se, dte);
      } else {
        orc_gh_time_derivative(n, ue, du, se, se + n, se + 2 * n, gauge_params,
                               se + 3 * n, se + 7 * n,
                               coords ? coords + (size_t)e * 3 * n : NULL, dte);
      }
      const double* ve = g_mesh_velocity ? g_mesh_velocity + (size_t)e * 3 * n : NULL;
      if (ve) {
        for (int p = 0; p < n; ++p) {
          const double v[3] = {ve[p], ve[(size_t)n + p], ve[(size_t)2 * n + p]};
          if (system == 1) {
            for (int s = 0; s < 10; ++s) {
              double vc3 = 0.0;
              for (int i = 0; i < 3; ++i)
                vc3 += v[i] * (du[(size_t)(3 * s + i) * n + p] - ue[(size_t)(20 + i + 3 * s) * n + p]);
              dte[(size_t)s * n + p] += se[(size_t)n + p] * vc3;
              dte[(size_t)(10 + s) * n + p] += se[(size_t)n + p] * se[(size_t)2 * n + p] * vc3;
            }
          }
          for (int c = 0; c < C; ++c) {
            double t = 0.0;
            for (int i = 0; i < 3; ++i) t += v[i] * du[(size_t)(3 * c + i) * n + p];
            dte[(size_t)c * n + p] += t;
          }
        }
      }
      /* faces: slice, normal, package */
      for (int d = 0; d < 6; ++d) {
        const int dim = d / 2;
        const double sign = (d % 2) ? 1.0 : -1.0;
        double* pk = pk_all + ((size_t)e * 6 + d) * PK * f;
        double* mag = mag_all + ((size_t)e * 6 + d) * f;
        for (int b = 0; b < N; ++b)
          for (int a = 0; a < N; ++a) {
            const int q = a + N * b;
            const int p = face_index(N, d, a, b);
            double unnorm[3], n_lo[3], n_up[3], up[50], out[134];
            for (int i = 0; i < 3; ++i)
              unnorm[i] = sign * je[(size_t)(dim + 3 * i) * n + p];
            for (int c = 0; c < C; ++c) up[c] = ue[(size_t)c * n + p];
            double ndotv = 0.0;
            if (system == 0) {
              face_normal(unnorm, 0, NULL, n_lo, n_up, &mag[q]);
              if (ve) for (int i = 0; i < 3; ++i) ndotv += n_lo[i] * ve[(size_t)i * n + p];
              sw_package_point_moving(up, se[p], n_lo, ndotv, out);
            } else {
              double g[4][4];
              GhGeom qg;
              for (int a4 = 0; a4 < 4; ++a4)
                for (int b4 = 0; b4 < 4; ++b4) g[a4][b4] = up[SYM4(a4, b4)];
              gh_geometry(g, &qg);
              face_normal(unnorm, 1, qg.inv_gamma, n_lo, n_up, &mag[q]);
              if (ve) for (int i = 0; i < 3; ++i) ndotv += n_lo[i] * ve[(size_t)i * n + p];
              gh_package_point_moving(up, se[n + p], se[2 * n + p], qg.lapse, qg.shift,
                                      n_lo, n_up, ndotv, out);
            }
            for (int c = 0; c < PK; ++c) pk[(size_t)c * f + q] = out[c];
          }
      }
    }
    free(du);
#pragma omp barrier
#pragma omp for schedule(static)
    for (int e = 0; e < nelem; ++e) {
      double* dte = dt_u + (size_t)e * C * n;
      for (int d = 0; d < 6; ++d) {
        const int ne = nbr[e * 6 + d];
        if (ne == -1 || ne == ORC_HANGING || ne == ORC_BJORHUS || ne == ORC_BJORHUS_PHYSICAL ||
            (ne < -1 && ext_u == NULL))
          continue;
        /* neighbour's face pointing back at us, and the face-point permutation */
        const int nf = (nbr_face && ne >= 0) ? nbr_face[e * 6 + d] : (d ^ 1);
        const int dn = nf & 7, perm = nf >> 3;
        const double* pki = pk_all + ((size_t)e * 6 + d) * PK * f;
        const double* pke = ne >= 0 ? pk_all + ((size_t)ne * 6 + dn) * PK * f : NULL;
        const double* mag = mag_all + ((size_t)e * 6 + d) * f;
        const double* ue = u + (size_t)e * C * n;
        const double* je = invjac + (size_t)e * 9 * n;
        const double* se = static_fields + (size_t)e * nstatic * n;
        for (int b = 0; b < N; ++b)
          for (int a = 0; a < N; ++a) {
            const int q = a + N * b;
            const int p = face_index(N, d, a, b);
            double in[134], ex[134], corr[50];
            for (int c = 0; c < PK; ++c) in[c] = pki[(size_t)c * f + q];
            if (ne >= 0) {
              int na = (perm & 1) ? b : a, nb2 = (perm & 1) ? a : b;
              if (perm & 2) na = N - 1 - na;
              if (perm & 4) nb2 = N - 1 - nb2;
              const int qn = na + N * nb2;
              for (int c = 0; c < PK; ++c) ex[c] = pke[(size_t)c * f + qn];
            } else {
              /* ghost boundary condition: package the exterior state */
              const double* xu = ext_u + (size_t)(-(ne + 2)) * C * f;
              double uex[50], unnorm[3], n_lo_i[3], n_up_i[3], mag_i, neg[3], n_lo[3],
                  n_up[3], mag_e;
              const int dim = d / 2;
              const double sign = (d % 2) ? 1.0 : -1.0;
              for (int c = 0; c < C; ++c) uex[c] = xu[(size_t)c * f + q];
              for (int i = 0; i < 3; ++i) unnorm[i] = sign * je[(size_t)(dim + 3 * i) * n + p];
              const double* vg = g_mesh_velocity ? g_mesh_velocity + (size_t)e * 3 * n : NULL;
              double ndotv_e = 0.0;
              if (system == 0) {
                face_normal(unnorm, 0, NULL, n_lo_i, n_up_i, &mag_i);
                for (int i = 0; i < 3; ++i) n_lo[i] = -n_lo_i[i];
                if (vg) for (int i = 0; i < 3; ++i) ndotv_e += n_lo[i] * vg[(size_t)i * n + p];
                sw_package_point_moving(uex, se[p], n_lo, ndotv_e, ex);
              } else {
                double g[4][4], gi[4][4];
                GhGeom qe, qi;
                for (int a4 = 0; a4 < 4; ++a4)
                  for (int b4 = 0; b4 < 4; ++b4) {
                    g[a4][b4] = uex[SYM4(a4, b4)];
                    gi[a4][b4] = ue[(size_t)SYM4(a4, b4) * n + p];
                  }
                gh_geometry(gi, &qi);
                face_normal(unnorm, 1, qi.inv_gamma, n_lo_i, n_up_i, &mag_i);
                gh_geometry(g, &qe);
                for (int i = 0; i < 3; ++i) neg[i] = -n_lo_i[i];
                face_normal(neg, 1, qe.inv_gamma, n_lo, n_up, &mag_e);
                if (vg) for (int i = 0; i < 3; ++i) ndotv_e += n_lo[i] * vg[(size_t)i * n + p];
                gh_package_point_moving(uex, se[n + p], se[2 * n + p], qe.lapse, qe.shift, n_lo,
                                        n_up, ndotv_e, ex);
              }
            }
            if (system == 0)
              sw_boundary_terms_point(in, ex, corr);
            else
              gh_boundary_terms_point(in, ex, corr);
            /* LiftFlux.hpp:57-61 */
            const double lift = -0.5 * (double)(N * (N - 1)) * mag[q];
            for (int c = 0; c < C; ++c) dte[(size_t)c * n + p] += corr[c] * lift;
          }
      }
    }
  }
  /* non-conforming mortars (serial: several mortars add to one coarse face) */
  for (int m = 0; m < n_mortars; ++m) {
    /* row[3] = fine direction | (perm << 3): perm takes a mortar point (a, b) in the
     * coarse element's face frame to the fine element's face point, with the bits of
     * nbr_face (orient_variables_on_slice of the received mortar data,
     * ApplyBoundaryCorrections.hpp:236-262 / OrientationMapHelpers.cpp:25-120) */
    const int ec = mortars[6 * m], dc = mortars[6 * m + 1], ef = mortars[6 * m + 2],
              df = mortars[6 * m + 3] & 7, permF = mortars[6 * m + 3] >> 3,
              sa = mortars[6 * m + 4], sb = mortars[6 * m + 5];
    const double* Pa = P + (size_t)sa * N * N;
    const double* Pb = P + (size_t)sb * N * N;
    const double* Ra = R + (size_t)sa * N * N;
    const double* Rb = R + (size_t)sb * N * N;
    const double* pkC = pk_all + ((size_t)ec * 6 + dc) * PK * f;
    const double* pkF = pk_all + ((size_t)ef * 6 + df) * PK * f;
    const double* magC = mag_all + ((size_t)ec * 6 + dc) * f;
    const double* magF = mag_all + ((size_t)ef * 6 + df) * f;
    double* pkCm = (double*)malloc(sizeof(double) * (size_t)PK * f);
    double* tmp = (double*)malloc(sizeof(double) * (size_t)f);
    double* corrC = (double*)malloc(sizeof(double) * (size_t)C * f);
    /* project_to_mortar: apply_matrices, first face dimension then the second */
    for (int c = 0; c < PK; ++c) {
      for (int b = 0; b < N; ++b)
        for (int a2 = 0; a2 < N; ++a2) {
          double v = 0.0;
          for (int a = 0; a < N; ++a) v += Pa[a2 * N + a] * pkC[(size_t)c * f + a + N * b];
          tmp[a2 + N * b] = v;
        }
      for (int b2 = 0; b2 < N; ++b2)
        for (int a2 = 0; a2 < N; ++a2) {
          double v = 0.0;
          for (int b = 0; b < N; ++b) v += Pb[b2 * N + b] * tmp[a2 + N * b];
          pkCm[(size_t)c * f + a2 + N * b2] = v;
        }
    }
    double* dtF = dt_u + (size_t)ef * C * n;
    double* dtC = dt_u + (size_t)ec * C * n;
    for (int b = 0; b < N; ++b)
      for (int a = 0; a < N; ++a) {
        const int q = a + N * b;
        int fa = (permF & 1) ? b : a, fb = (permF & 1) ? a : b;
        if (permF & 2) fa = N - 1 - fa;
        if (permF & 4) fb = N - 1 - fb;
        const int qF = fa + N * fb;
        double pc[134], pf[134], corr[50];
        for (int c = 0; c < PK; ++c) {
          pc[c] = pkCm[(size_t)c * f + q];
          pf[c] = pkF[(size_t)c * f + qF];
        }
        /* the fine element: its face is the mortar */
        if (system == 0)
          sw_boundary_terms_point(pf, pc, corr);
        else
          gh_boundary_terms_point(pf, pc, corr);
        const int p = face_index(N, df, fa, fb);
        const double lift = -0.5 * (double)(N * (N - 1)) * magF[qF];
        for (int c = 0; c < C; ++c) dtF[(size_t)c * n + p] += corr[c] * lift;
        /* the coarse element's correction on the mortar */
        if (system == 0)
          sw_boundary_terms_point(pc, pf, corr);
        else
          gh_boundary_terms_point(pc, pf, corr);
        for (int c = 0; c < C; ++c) corrC[(size_t)c * f + q] = corr[c];
      }
    /* project_from_mortar, lift on the coarse face, add_slice_to_data */
    for (int c = 0; c < C; ++c) {
      for (int b2 = 0; b2 < N; ++b2)
        for (int a = 0; a < N; ++a) {
          double v = 0.0;
          for (int a2 = 0; a2 < N; ++a2) v += Ra[a * N + a2] * corrC[(size_t)c * f + a2 + N * b2];
          tmp[a + N * b2] = v;
        }
      for (int b = 0; b < N; ++b)
        for (int a = 0; a < N; ++a) {
          double v = 0.0;
          for (int b2 = 0; b2 < N; ++b2) v += Rb[b * N + b2] * tmp[a + N * b2];
          const int q = a + N * b;
          const int p = face_index(N, dc, a, b);
          dtC[(size_t)c * n + p] += v * (-0.5 * (double)(N * (N - 1)) * magC[q]);
        }
    }
    free(pkCm);
    free(tmp);
    free(corrC);
  }
  free(pk_all);
  free(mag_all);
}

/* volume-only part (no faces), for kernel-level parity tests */
void orc_dg_volume(int system, int N, int nelem, const double* D,
                   const double* u, const double* invjac,
                   const double* static_fields, const double* coords,
                   const double* gauge_params, double* dt_u) {
  const int n = N * N * N;
  const int C = system == 0 ? 5 : 50;
  GhGaugeSpec gs;
  gs.gauge_mode = 0;
  if (system == 1) gauge_from_params(gauge_params, &gs);
  const int nstatic = system == 0 ? 1 : (gs.gauge_mode == 1 ? 23 : 3);
#pragma omp parallel
  {
    double* du = (double*)malloc(sizeof(double) * 3 * (size_t)C * n);
#pragma omp for schedule(static)
    for (int e = 0; e < nelem; ++e) {
      const double* ue = u + (size_t)e * C * n;
      const double* je = invjac + (size_t)e * 9 * n;
      const double* se = static_fields + (size_t)e * nstatic * n;
      double* dte = dt_u + (size_t)e * C * n;
      orc_partial_derivatives(N, C, D, ue, je, du);
      if (system == 0)
        orc_sw_time_derivative(n, ue, du, se, dte);
      else
        orc_gh_time_derivative(n, ue, du, se, se + n, se + 2 * n, gauge_params,
                               se + 3 * n, se + 7 * n,
                               coords ? coords + (size_t)e * 3 * n : NULL, dte);
    }
    free(du);
  }
}

/* u <- a*u + sum_j c_j * v_j  (flat axpy chain of the steppers,
 * AdamsBashforth.cpp:176-201, RungeKutta.cpp:69-82) */
void orc_lincomb(long long len, double a, double* u, int nterms,
                 const double* coefs, const double* const* vs) {
#pragma omp parallel for schedule(static)
  for (long long p = 0; p < len; ++p) {
    double v = a * u[p];
    for (int j = 0; j < nterms; ++j) v += coefs[j] * vs[j][p];
    u[p] = v;
  }
}

/* Exponential filter applied to every component block: apply_matrices(u, {F, F, F})
 * (NumericalAlgorithms/LinearOperators/ExponentialFilter.cpp:45-76 ->
 * ApplyMatrices.hpp), F row-major [N][N]; u holds `nblocks` blocks of N^3
 * doubles (xi fastest), filtered in place, one dimension after the other. */
void orc_apply_filter(int N, long long nblocks, const double* F, double* u) {
  const int n = N * N * N;
#pragma omp parallel
  {
    double* a = (double*)malloc(sizeof(double) * (size_t)n);
    double* b = (double*)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(static)
    for (long long blk = 0; blk < nblocks; ++blk) {
      double* v = u + blk * n;
      for (int k = 0; k < N; ++k)
        for (int j = 0; j < N; ++j)
          for (int i = 0; i < N; ++i) {
            double s = 0.0;
            for (int m = 0; m < N; ++m) s += F[i * N + m] * v[m + N * (j + N * k)];
            a[i + N * (j + N * k)] = s;
          }
      for (int k = 0; k < N; ++k)
        for (int j = 0; j < N; ++j)
          for (int i = 0; i < N; ++i) {
            double s = 0.0;
            for (int m = 0; m < N; ++m) s += F[j * N + m] * a[i + N * (m + N * k)];
            b[i + N * (j + N * k)] = s;
          }
      for (int k = 0; k < N; ++k)
        for (int j = 0; j < N; ++j)
          for (int i = 0; i < N; ++i) {
            double s = 0.0;
            for (int m = 0; m < N; ++m) s += F[k * N + m] * b[i + N * (j + N * m)];
            v[i + N * (j + N * k)] = s;
          }
    }
    free(a);
    free(b);
  }
}

/* ------------------------------------------------------------------------
 * DampedHarmonic gauge without roll-on: damped_harmonic_impl<false>
 * Evolution/Systems/GeneralizedHarmonic/GaugeSourceFunctions/
 *   DampedHarmonic.cpp:70-439, DampedWaveHelpers.cpp:26-62
 * (spatial_weight_function W = exp(-r^2/sigma_r^2), its spacetime derivative,
 * log_factor_metric_lapse).  Pinned by fixtures made with the reference's
 * DampedHarmonic.py (tests/golden/damped_harmonic.npz).
 * d4_g[a][b][c] = d_a g_bc; d4H[a][b] = d_a H_b.
 * ---------------------------------------------------------------------- */
static double integer_pow(double x, int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= x;
  return r;
}

static void damped_harmonic_gauge(
    const DampedHarmonicParams* prm, const double x[3], double lapse,
    const double shift[3], double sqrt_det_gamma, const double inv_gamma[3][3],
    const double d4_g[4][4][4], double half_pi_two_normals,
    const double half_phi_two_normals[3], const double g[4][4],
    const double phi[3][4][4], double H[4], double d4H[4][4]) {
  const double amp_L1 = prm->amp[0], amp_L2 = prm->amp[1], amp_S = prm->amp[2];
  const int exp_L1 = prm->exp[0], exp_L2 = prm->exp[1], exp_S = prm->exp[2];
  const double sigma_r = prm->width;
  const double one_over_lapse = 1.0 / lapse;
  const double log_fac_1 = log(sqrt_det_gamma / lapse); /* exponent 0.5 */
  const double log_fac_2 = -log(lapse);                 /* exponent 0   */
  /* DampedWaveHelpers.cpp:26-47 */
  const double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  const double weight = exp(-r2 / (sigma_r * sigma_r));
  double d4_weight[4];
  d4_weight[0] = 0.0;
  for (int i = 0; i < 3; ++i)
    d4_weight[i + 1] = -2.0 * weight / (sigma_r * sigma_r) * x[i];
  double pow1 = integer_pow(log_fac_1, exp_L1);
  double pow2 = integer_pow(log_fac_1, exp_S);
  double pow3 = integer_pow(log_fac_2, exp_L2);
  const double mu_L1 = amp_L1 * weight * pow1;
  const double mu_S = amp_S * weight * pow2;
  const double mu_L2 = amp_L2 * weight * pow3;
  const double mu_S_over_lapse = mu_S * one_over_lapse;
  const double mu1 = mu_L1 * log_fac_1;
  const double mu2 = mu_L2 * log_fac_2;
  const double prefac = mu_L1 * log_fac_1 + mu_L2 * log_fac_2;
  double g_dot_shift[4];
  for (int a = 0; a < 4; ++a) {
    g_dot_shift[a] = g[a][1] * shift[0];
    for (int i = 1; i < 3; ++i) g_dot_shift[a] += g[a][i + 1] * shift[i];
  }
  for (int a = 0; a < 4; ++a) H[a] = -mu_S_over_lapse * g_dot_shift[a];
  H[0] -= prefac * lapse;

  /* d_t lapse and d_a lapse / lapse (:262-270) */
  double sh_hphi = 0.0;
  for (int i = 0; i < 3; ++i) sh_hphi += shift[i] * half_phi_two_normals[i];
  const double dt_lapse = lapse * (lapse * half_pi_two_normals - sh_hphi);
  double d_lapse_by_lapse[4];
  d_lapse_by_lapse[0] = one_over_lapse * dt_lapse;
  for (int i = 0; i < 3; ++i) d_lapse_by_lapse[i + 1] = -half_phi_two_normals[i];
  /* d_a det(gamma) / det(gamma) = gamma^{jk} d_a gamma_jk (:271-277) */
  double d_g_by_det[4];
  for (int a = 0; a < 4; ++a) {
    double v = 0.0;
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) v += inv_gamma[j][k] * d4_g[a][j + 1][k + 1];
    d_g_by_det[a] = v;
  }
  double d4_log_fac_mu1[4], d4_log_fac_muS[4], d4_log_fac_mu2[4];
  for (int a = 0; a < 4; ++a) {
    const double d_logfac_1 = 0.5 * d_g_by_det[a] - d_lapse_by_lapse[a];
    const double d_logfac_2 = -d_lapse_by_lapse[a];
    d4_log_fac_mu1[a] = (double)(exp_L1 + 1) * integer_pow(log_fac_1, exp_L1) * d_logfac_1;
    d4_log_fac_muS[a] = (double)exp_S * integer_pow(log_fac_1, exp_S - 1) * d_logfac_1;
    d4_log_fac_mu2[a] = (double)(exp_L2 + 1) * integer_pow(log_fac_2, exp_L2) * d_logfac_2;
  }
  pow1 *= log_fac_1 * amp_L1;
  pow2 *= amp_S;
  pow3 *= log_fac_2 * amp_L2;
  double d4_mu1[4], d4_mu_S[4], d4_mu2[4];
  for (int a = 0; a < 4; ++a) {
    d4_mu1[a] = pow1 * d4_weight[a] + amp_L1 * weight * d4_log_fac_mu1[a];
    d4_mu_S[a] = d4_weight[a] * pow2 + amp_S * weight * d4_log_fac_muS[a];
    d4_mu2[a] = pow3 * d4_weight[a] + amp_L2 * weight * d4_log_fac_mu2[a];
  }
  /* :359-379 */
  double d4_muS_over_lapse[4], dT2[4];
  d4_muS_over_lapse[0] = dt_lapse;
  dT2[0] = -(d4_mu1[0] + d4_mu2[0]) * lapse - (mu1 + mu2) * d4_muS_over_lapse[0];
  for (int i = 0; i < 3; ++i)
    dT2[i + 1] = -(d4_mu1[i + 1] + d4_mu2[i + 1]) * lapse +
                 (mu1 + mu2) * lapse * half_phi_two_normals[i];
  d4_muS_over_lapse[0] *= -mu_S * one_over_lapse;
  d4_muS_over_lapse[0] += d4_mu_S[0];
  d4_muS_over_lapse[0] *= one_over_lapse;
  for (int i = 0; i < 3; ++i)
    d4_muS_over_lapse[i + 1] =
        one_over_lapse * (d4_mu_S[i + 1] + mu_S * half_phi_two_normals[i]);
  /* :397-421 */
  for (int a = 0; a < 4; ++a) {
    double dT3[4];
    for (int j = 0; j < 3; ++j) dT3[j + 1] = d4_g[a][0][j + 1];
    dT3[0] = d4_g[a][0][1] * shift[0];
    for (int j = 1; j < 3; ++j) dT3[0] += d4_g[a][0][j + 1] * shift[j];
    for (int i = 0; i < 3; ++i)
      for (int j = i + 1; j < 3; ++j)
        dT3[0] -= shift[i] * shift[j] * d4_g[a][i + 1][j + 1];
    dT3[0] *= 2.0;
    for (int i = 0; i < 3; ++i) dT3[0] -= shift[i] * shift[i] * d4_g[a][i + 1][i + 1];
    for (int b = 0; b < 4; ++b) {
      dT3[b] *= -mu_S_over_lapse;
      dT3[b] -= d4_muS_over_lapse[a] * g_dot_shift[b];
      d4H[a][b] = dT3[b];
    }
    d4H[a][0] += dT2[a];
  }
  (void)phi;
}

/* stand-alone evaluation for the pin test: g, pi, phi dense arrays at a point */
void orc_damped_harmonic(const double* g_, const double* pi_, const double* phi_,
                         const double* x, const double* gauge_params, double* H,
                         double* d4H_) {
  const double(*g)[4] = (const double(*)[4])g_;
  const double(*pi)[4] = (const double(*)[4])pi_;
  const double(*phi)[4][4] = (const double(*)[4][4])phi_;
  double(*d4H)[4] = (double(*)[4])d4H_;
  GhGaugeSpec gs;
  gauge_from_params(gauge_params, &gs);
  GhGeom q;
  gh_geometry(g, &q);
  double da_g[4][4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double v = -q.lapse * pi[a][b];
      for (int m = 0; m < 3; ++m) v += q.shift[m] * phi[m][a][b];
      da_g[0][a][b] = v;
      for (int i = 0; i < 3; ++i) da_g[i + 1][a][b] = phi[i][a][b];
    }
  double pon[4], hpnn = 0.0, hphinn[3];
  for (int a = 0; a < 4; ++a) {
    pon[a] = 0.0;
    for (int b = 0; b < 4; ++b) pon[a] += q.normal_vec[b] * pi[b][a];
  }
  for (int a = 0; a < 4; ++a) hpnn += q.normal_vec[a] * pon[a];
  hpnn *= 0.5;
  for (int n = 0; n < 3; ++n) {
    double v = 0.0;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) v += q.normal_vec[a] * q.normal_vec[b] * phi[n][a][b];
    hphinn[n] = 0.5 * v;
  }
  damped_harmonic_gauge(&gs.dh, x, q.lapse, q.shift, sqrt(q.det_gamma), q.inv_gamma, da_g,
                        hpnn, hphinn, g, phi, H, d4H);
}
