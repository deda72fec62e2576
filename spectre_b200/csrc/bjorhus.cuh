// Pointwise algebra of gh::BoundaryConditions::ConstraintPreservingBjorhus,
// Types ConstraintPreserving and ConstraintPreservingPhysical, static mesh
// (GeneralizedHarmonic/BoundaryConditions/
// Bjorhus.cpp:104-391, compute_intermediate_vars :393-545, BjorhusImpl.cpp:26-221,
// 496-532; constraints: Constraints.hpp Eq. (43)/(44) of Lindblom et al. 2005 as
// implemented in Constraints.cpp:25-280 and :282-1262; characteristic fields
// Characteristics.cpp:57-169).  __host__ __device__ like pointwise.cuh so that the
// CPU harness can compare it with the oracle.
//
// Tensors are full (unpacked) arrays: spacetime indices 0..3, spatial 0..2; a
// spatial index in a spacetime slot means index + 1.
#pragma once

#include "pointwise.cuh"

namespace dg {

struct BjorhusInput {
  double n_lo[3];          // outward unit normal covector of the face
  double g[4][4], pi[4][4], phi[3][4][4];
  double x[3];             // inertial coordinates (Sommerfeld radius)
  double gamma1, gamma2;
  double lapse, shift[3];
  double ipsi[4][4];       // inverse spacetime metric
  double t_up[4];          // spacetime unit normal vector
  double c3[3][4][4];      // three-index constraint d_i g_ab - Phi_iab
  double H[4], dH[4][4];   // gauge source H_a and d_a H_b
  double dt_g[4][4], dt_pi[4][4], dt_phi[3][4][4];  // volume time derivative
  double d_pi[3][4][4], d_phi[3][3][4][4];          // d_i Pi_ab, d_i Phi_jab
  bool physical;           // Type ConstraintPreservingPhysical (else ConstraintPreserving)
};

struct BjorhusOutput {
  double g[4][4], pi[4][4], phi[3][4][4];  // corrections ADDED to dt(g, Pi, Phi)
};

DG_HD double levi_civita(int i, int j, int k) {
  return 0.5 * (double)((i - j) * (j - k) * (k - i));
}

// deliberately NOT force-inlined: the function is large and independent of the
// number of grid points, one copy serves every kernel instantiation
#ifdef __CUDACC__
#define DG_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define DG_HD_NOINLINE static
#endif
DG_HD_NOINLINE void bjorhus_constraint_preserving(const BjorhusInput& in, BjorhusOutput& out) {
  // ---- geometry of the slice and of the face -------------------------------
  double t_lo[4] = {-in.lapse, 0.0, 0.0, 0.0};
  double ig[3][3], n_up[3];
  const double il2 = 1.0 / (in.lapse * in.lapse);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) ig[i][j] = in.ipsi[i + 1][j + 1] + in.shift[i] * in.shift[j] * il2;
  for (int i = 0; i < 3; ++i) {
    double v = 0.0;
    for (int j = 0; j < 3; ++j) v += ig[i][j] * in.n_lo[j];
    n_up[i] = v;
  }
  double sdn = 0.0;
  for (int i = 0; i < 3; ++i) sdn += in.shift[i] * in.n_lo[i];
  const double speed[4] = {-(1.0 + in.gamma1) * sdn, -sdn, -sdn + in.lapse, -sdn - in.lapse};
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) out.g[a][b] = out.pi[a][b] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) out.phi[i][a][b] = 0.0;
  double mn = speed[0];
  for (int k = 1; k < 4; ++k) mn = speed[k] < mn ? speed[k] : mn;
  if (mn >= 0.0) return;  // nothing enters the domain at this point

  // g_a^i = delta_a^i + t^i t_a
  double gm[4][3];
  for (int a = 0; a < 4; ++a)
    for (int i = 0; i < 3; ++i) gm[a][i] = (a == i + 1 ? 1.0 : 0.0) + in.t_up[i + 1] * t_lo[a];

  // ---- two-index constraint C_ia (Eq. 44) and F constraint (Eq. 43) ----------
  // frequently used contractions
  double phi_up[3][4][4];   // Phi_i^{cd}
  double tr_phi[3];         // psi^{cd} Phi_icd
  double phi_t[3][4];       // Phi_iab t^b
  double phi_tt[3];         // Phi_icd t^c t^d
  for (int i = 0; i < 3; ++i) {
    for (int c = 0; c < 4; ++c)
      for (int d = 0; d < 4; ++d) {
        double v = 0.0;
        for (int e = 0; e < 4; ++e) {
          double w = 0.0;
          for (int f = 0; f < 4; ++f) w += in.ipsi[d][f] * in.phi[i][e][f];
          v += in.ipsi[c][e] * w;
        }
        phi_up[i][c][d] = v;
      }
    double tr = 0.0, tt = 0.0;
    for (int c = 0; c < 4; ++c) {
      double pt = 0.0;
      for (int d = 0; d < 4; ++d) {
        tr += in.ipsi[c][d] * in.phi[i][c][d];
        pt += in.phi[i][c][d] * in.t_up[d];
      }
      phi_t[i][c] = pt;
      tt += pt * in.t_up[c];
    }
    tr_phi[i] = tr;
    phi_tt[i] = tt;
  }
  double pi_t[4], pi_tt = 0.0, tr_pi = 0.0;   // Pi_ab t^b, Pi_ab t^a t^b, psi^{ab} Pi_ab
  double pi_up[4][4];                         // Pi^{cd}... as psi^{cb} psi^{de} Pi_be
  for (int a = 0; a < 4; ++a) {
    double v = 0.0;
    for (int b = 0; b < 4; ++b) {
      v += in.pi[a][b] * in.t_up[b];
      tr_pi += in.ipsi[a][b] * in.pi[a][b];
    }
    pi_t[a] = v;
    pi_tt += v * in.t_up[a];
  }
  for (int c = 0; c < 4; ++c)
    for (int d = 0; d < 4; ++d) {
      double v = 0.0;
      for (int b = 0; b < 4; ++b) {
        double w = 0.0;
        for (int e = 0; e < 4; ++e) w += in.ipsi[d][e] * in.pi[b][e];
        v += in.ipsi[c][b] * w;
      }
      pi_up[c][d] = v;
    }
  double tr_dpi[3], tr_c3[3], tr_dphi[3][3];  // psi^{cd} d_i Pi_cd, psi^{cd} C_icd, psi^{cd} d_j Phi_icd
  for (int i = 0; i < 3; ++i) {
    double a1 = 0.0, a2 = 0.0;
    for (int c = 0; c < 4; ++c)
      for (int d = 0; d < 4; ++d) {
        a1 += in.ipsi[c][d] * in.d_pi[i][c][d];
        a2 += in.ipsi[c][d] * in.c3[i][c][d];
      }
    tr_dpi[i] = a1;
    tr_c3[i] = a2;
    for (int j = 0; j < 3; ++j) {
      double v = 0.0;
      for (int c = 0; c < 4; ++c)
        for (int d = 0; d < 4; ++d) v += in.ipsi[c][d] * in.d_phi[j][i][c][d];
      tr_dphi[j][i] = v;
    }
  }

  // gg_phi[k][n][a] = g^{jk} g^{mn} Phi_jma
  double gg_phi[3][3][4];
  for (int k = 0; k < 3; ++k)
    for (int n = 0; n < 3; ++n)
      for (int a = 0; a < 4; ++a) {
        double v = 0.0;
        for (int j = 0; j < 3; ++j) {
          double w = 0.0;
          for (int m = 0; m < 3; ++m) w += ig[m][n] * in.phi[j][m + 1][a];
          v += ig[j][k] * w;
        }
        gg_phi[k][n][a] = v;
      }
  double c2[3][4];
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 4; ++a) {
      double v = 0.0;
      // g^{jk} d_j Phi_ika
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) v += ig[j][k] * in.d_phi[j][i][k + 1][a];
      // -1/2 g_a^j psi^{cd} d_j Phi_icd
      for (int j = 0; j < 3; ++j) v -= 0.5 * gm[a][j] * tr_dphi[j][i];
      // t^b d_i Pi_ba - 1/2 t_a psi^{cd} d_i Pi_cd + d_i H_a
      for (int b = 0; b < 4; ++b) v += in.t_up[b] * in.d_pi[i][b][a];
      v -= 0.5 * t_lo[a] * tr_dpi[i];
      v += in.dH[i + 1][a];
      // 1/2 g_a^j Phi_jcd Phi_i^{cd}
      for (int j = 0; j < 3; ++j) {
        double w = 0.0;
        for (int c = 0; c < 4; ++c)
          for (int d = 0; d < 4; ++d) w += in.phi[j][c][d] * phi_up[i][c][d];
        v += 0.5 * gm[a][j] * w;
      }
      // 1/2 g^{jk} (psi^{cd} Phi_jcd) Phi_ike t^e t_a
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) v += 0.5 * ig[j][k] * tr_phi[j] * phi_t[i][k + 1] * t_lo[a];
      // - g^{jk} g^{mn} Phi_jma Phi_ikn = - X[k][n][a] Y[i][k][n]
      for (int k = 0; k < 3; ++k)
        for (int n = 0; n < 3; ++n) v -= gg_phi[k][n][a] * in.phi[i][k + 1][n + 1];
      // 1/2 Phi_icd Pi_be t_a (psi^{cb} psi^{de} + 1/2 psi^{be} t^c t^d)
      {
        double w = 0.0;
        for (int c = 0; c < 4; ++c)
          for (int d = 0; d < 4; ++d) w += in.phi[i][c][d] * pi_up[c][d];
        v += 0.5 * t_lo[a] * (w + 0.5 * tr_pi * phi_tt[i]);
      }
      // - Phi_icd Pi_ba t^c (psi^{bd} + 1/2 t^b t^d)
      {
        double w = 0.0;
        for (int b = 0; b < 4; ++b)
          for (int d = 0; d < 4; ++d) w += phi_t[i][d] * in.pi[b][a] * in.ipsi[b][d];
        v -= w + 0.5 * phi_tt[i] * pi_t[a];
      }
      // 1/2 gamma2 t_a psi^{cd} C_icd - gamma2 t^d C_iad
      v += 0.5 * in.gamma2 * t_lo[a] * tr_c3[i];
      for (int d = 0; d < 4; ++d) v -= in.gamma2 * in.t_up[d] * in.c3[i][a][d];
      c2[i][a] = v;
    }

  // F_a
  double fc[4];
  {
    // scalars multiplying t_a
    double s_ta = 0.0;
    double tr_dphi_sp = 0.0, div_H = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        tr_dphi_sp += ig[i][j] * tr_dphi[i][j];          // psi^{bc} g^{ij} d_i Phi_jbc
        div_H += ig[i][j] * in.dH[i + 1][j + 1];         // g^{ij} d_i H_j
      }
    s_ta += 0.5 * tr_dphi_sp + div_H;
    // -1/2 g^{ij} g^{mn} Phi_imc Phi_njd psi^{cd}, with U[i][n][c] = g^{mn} Phi_imc and
    // W[n][i][c] = psi^{cd} g^{ij} Phi_njd
    {
      double U[3][3][4], W[3][3][4];
      for (int i = 0; i < 3; ++i)
        for (int n = 0; n < 3; ++n)
          for (int c = 0; c < 4; ++c) {
            double u = 0.0, r = 0.0;
            for (int m = 0; m < 3; ++m) {
              u += ig[m][n] * in.phi[i][m + 1][c];
              r += ig[n][m] * in.phi[i][m + 1][c];   // g^{nj} Phi_{i j c} (row n of g^{..})
            }
            U[i][n][c] = u;
            W[i][n][c] = r;                          // raised with psi below
          }
      double w = 0.0;
      for (int i = 0; i < 3; ++i)
        for (int n = 0; n < 3; ++n)
          for (int c = 0; c < 4; ++c) {
            // psi^{cd} g^{ij} Phi_{n j d}: W[n][i][d] contracted with psi
            double r = 0.0;
            for (int d = 0; d < 4; ++d) r += in.ipsi[c][d] * W[n][i][d];
            w += U[i][n][c] * r;
          }
      s_ta -= 0.5 * w;
    }
    // -1/4 g^{ij} Phi_icd Phi_j^{cd}
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double w = 0.0;
        for (int c = 0; c < 4; ++c)
          for (int d = 0; d < 4; ++d) w += in.phi[i][c][d] * phi_up[j][c][d];
        s_ta -= 0.25 * ig[i][j] * w;
      }
    // +1/4 Pi_cd Pi^{cd} + 1/2 Pi_cd Pi_be psi^{ce} t^d t^b
    {
      double w = 0.0, w2 = 0.0;
      for (int c = 0; c < 4; ++c)
        for (int d = 0; d < 4; ++d) {
          w += in.pi[c][d] * pi_up[c][d];
          w2 += pi_t[c] * pi_t[d] * in.ipsi[c][d];
        }
      s_ta += 0.25 * w + 0.5 * w2;
    }
    // +1/2 (Pi_cd psi^{cd}) H_b t^b - g^{ij} Phi_ijc H_d psi^{cd} + 1/2 g^{ij} H_i Phi_jcd psi^{cd}
    {
      double Ht = 0.0;
      for (int b = 0; b < 4; ++b) Ht += in.H[b] * in.t_up[b];
      s_ta += 0.5 * tr_pi * Ht;
      double H_up[4];
      for (int c = 0; c < 4; ++c) {
        double w = 0.0;
        for (int d = 0; d < 4; ++d) w += in.ipsi[c][d] * in.H[d];
        H_up[c] = w;
      }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double w = 0.0;
          for (int c = 0; c < 4; ++c) w += in.phi[i][j + 1][c] * H_up[c];
          s_ta -= ig[i][j] * w;
          s_ta += 0.5 * ig[i][j] * in.H[i + 1] * tr_phi[j];
        }
    }
    // vectors contracted with g_a^i
    double vg[3];
    for (int i = 0; i < 3; ++i) {
      double v = 0.5 * tr_dpi[i];
      // Phi_ijb g^{jk} Phi_kcd psi^{bd} t^c - 1/2 Phi_ijb g^{jk} (psi^{cd} Phi_kcd) t^b
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) {
          double w = 0.0;
          for (int b = 0; b < 4; ++b)
            for (int d = 0; d < 4; ++d) w += in.phi[i][j + 1][b] * in.ipsi[b][d] * phi_t[k][d];
          v += ig[j][k] * (w - 0.5 * phi_t[i][j + 1] * tr_phi[k]);
        }
      // - t^b d_i H_b
      for (int b = 0; b < 4; ++b) v -= in.t_up[b] * in.dH[i + 1][b];
      // -1/4 Phi_icd t^c t^d Pi_be psi^{be} + Phi_icd Pi_be t^c t^b psi^{de}
      v -= 0.25 * phi_tt[i] * tr_pi;
      for (int d = 0; d < 4; ++d)
        for (int e = 0; e < 4; ++e) v += phi_t[i][d] * pi_t[e] * in.ipsi[d][e];
      // + Phi_icd H_b psi^{bc} t^d
      for (int b = 0; b < 4; ++b)
        for (int c = 0; c < 4; ++c) v += phi_t[i][c] * in.H[b] * in.ipsi[b][c];
      // -1/2 gamma2 psi^{cd} C_icd
      v -= 0.5 * in.gamma2 * tr_c3[i];
      vg[i] = v;
    }
    for (int a = 0; a < 4; ++a) {
      double v = s_ta * t_lo[a];
      for (int i = 0; i < 3; ++i) v += gm[a][i] * vg[i];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          // - g^{ij} d_i Pi_ja - g^{ij} t^b d_i Phi_jba
          v -= ig[i][j] * in.d_pi[i][j + 1][a];
          for (int b = 0; b < 4; ++b) v -= ig[i][j] * in.t_up[b] * in.d_phi[i][j][b][a];
          // + g^{ij} Phi_icd Phi_jba psi^{bc} t^d
          for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c) v += ig[i][j] * phi_t[i][c] * in.phi[j][b][a] * in.ipsi[b][c];
          // - g^{ij} H_i Pi_ja - t^b g^{ij} Pi_bi Pi_ja
          v -= ig[i][j] * in.H[i + 1] * in.pi[j + 1][a];
          v -= ig[i][j] * pi_t[i + 1] * in.pi[j + 1][a];
          // - g^{ij} Phi_iba t^b Pi_je t^e - 1/2 g^{ij} Phi_icd t^c t^d Pi_ja
          v -= ig[i][j] * phi_t[i][a] * pi_t[j + 1];
          v -= 0.5 * ig[i][j] * phi_tt[i] * in.pi[j + 1][a];
          // - g^{ij} H_i Phi_jba t^b
          v -= ig[i][j] * in.H[i + 1] * phi_t[j][a];
        }
      // gamma2 g^{id} C_ida with g^{id} = psi^{id} + t^i t^d (d over all four values)
      for (int i = 0; i < 3; ++i)
        for (int d = 0; d < 4; ++d)
          v += in.gamma2 * (in.ipsi[i + 1][d] + in.t_up[i + 1] * in.t_up[d]) * in.c3[i][d][a];
      fc[a] = v;
    }
  }

  // ---- characteristic projections of the volume time derivative ---------------
  double rhs_plus[4][4], rhs_minus[4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double pn = 0.0;
      for (int i = 0; i < 3; ++i) pn += n_up[i] * in.dt_phi[i][a][b];
      rhs_plus[a][b] = in.dt_pi[a][b] + pn - in.gamma2 * in.dt_g[a][b];
      rhs_minus[a][b] = in.dt_pi[a][b] - pn - in.gamma2 * in.dt_g[a][b];
    }
  // null normals, projectors
  const double r2 = sqrt(0.5);
  double in_lo[4], out_lo[4], in_up[4], out_up[4];
  for (int a = 0; a < 4; ++a) {
    const double nl = a == 0 ? 0.0 : in.n_lo[a - 1], nu = a == 0 ? 0.0 : n_up[a - 1];
    in_lo[a] = r2 * (t_lo[a] - nl);
    out_lo[a] = r2 * (t_lo[a] + nl);
    in_up[a] = r2 * (in.t_up[a] - nu);
    out_up[a] = r2 * (in.t_up[a] + nu);
  }
  double p_lo[4][4], p_mix[4][4], p_up[4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      const double nla = a == 0 ? 0.0 : in.n_lo[a - 1], nlb = b == 0 ? 0.0 : in.n_lo[b - 1];
      const double nua = a == 0 ? 0.0 : n_up[a - 1], nub = b == 0 ? 0.0 : n_up[b - 1];
      p_lo[a][b] = in.g[a][b] + t_lo[a] * t_lo[b] - nla * nlb;
      p_mix[a][b] = (a == b ? 1.0 : 0.0) + in.t_up[a] * t_lo[b] - nua * nlb;
      p_up[a][b] = in.ipsi[a][b] + in.t_up[a] * in.t_up[b] - nua * nub;
    }

  // ---- corrections to the characteristic fields --------------------------------
  double bc_psi[4][4], bc_zero[3][4][4], bc_plus[4][4], bc_minus[4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double v = 0.0;
      for (int i = 0; i < 3; ++i) v += n_up[i] * in.c3[i][a][b];
      bc_psi[a][b] = speed[0] * v;                     // BjorhusImpl.cpp:26-47
      bc_plus[a][b] = -rhs_plus[a][b];                 // Bjorhus.cpp:317-325
    }
  {
    // four-index constraint C_jab = eps_{j l m} d_l Phi_mab (Constraints.cpp:1070-1100)
    double c4[3][4][4];
    for (int j = 0; j < 3; ++j) {
      const int l = (j + 1) % 3, m = (j + 2) % 3;
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) c4[j][a][b] = in.d_phi[l][m][a][b] - in.d_phi[m][l][a][b];
    }
    // dt v_zero_iab = lambda_0 eps_{i j k} n^k C_jab (BjorhusImpl.cpp:49-103)
    for (int i = 0; i < 3; ++i) {
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
          bc_zero[i][a][b] = speed[1] * (n_up[k] * c4[j][a][b] - n_up[j] * c4[k][a][b]);
    }
  }
  {
    // constraint-dependent terms (BjorhusImpl.cpp:153-221, mu = 0) and gauge
    // Sommerfeld terms (:105-151)
    double common[4];
    for (int a = 0; a < 4; ++a) {
      double v = 0.0;
      for (int i = 0; i < 3; ++i) v += n_up[i] * c2[i][a];
      common[a] = r2 * speed[3] * (fc[a] + v);          // c^{0-}_a = F_a + n^k C_ka
    }
    double uAu = 0.0, trA = 0.0, in_A[4], A_in[4];
    for (int c = 0; c < 4; ++c) {
      double v1 = 0.0, v2 = 0.0;
      for (int d = 0; d < 4; ++d) {
        v1 += in_up[d] * rhs_minus[d][c];   // u^d A_dc
        v2 += rhs_minus[c][d] * in_up[d];   // A_cd u^d
        trA += p_up[c][d] * rhs_minus[c][d];
      }
      in_A[c] = v1;
      A_in[c] = v2;
      uAu += in_up[c] * v2;
    }
    double pinA[4], pAin[4], pc[4], pB[4];
    double u_common = 0.0, v_common = 0.0, uBv = 0.0, vBv = 0.0;
    for (int a = 0; a < 4; ++a) {
      double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
      for (int c = 0; c < 4; ++c) {
        s1 += p_mix[c][a] * in_A[c];
        s2 += p_mix[c][a] * A_in[c];
        s3 += p_mix[c][a] * common[c];
        double bd = 0.0;
        for (int d = 0; d < 4; ++d) bd += in.dt_g[c][d] * out_up[d];   // B_cd v^d
        s4 += p_mix[c][a] * bd;
      }
      pinA[a] = s1;
      pAin[a] = s2;
      pc[a] = s3;
      pB[a] = s4;
      u_common += in_up[a] * common[a];
      v_common += out_up[a] * common[a];
      double bd = 0.0;
      for (int d = 0; d < 4; ++d) bd += in.dt_g[a][d] * out_up[d];
      uBv += in_up[a] * bd;
      vBv += out_up[a] * bd;
    }
    const double radius = sqrt(in.x[0] * in.x[0] + in.x[1] * in.x[1] + in.x[2] * in.x[2]);
    const double prefac = in.gamma2 - 1.0 / radius;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        double v = 0.5 * (2.0 * uAu * out_lo[a] * out_lo[b] - pinA[a] * out_lo[b] -
                          out_lo[a] * pinA[b] - pAin[a] * out_lo[b] - out_lo[a] * pAin[b] +
                          trA * p_lo[a][b]);
        v += u_common * out_lo[a] * out_lo[b] + v_common * p_lo[a][b] - out_lo[a] * pc[b] -
             pc[a] * out_lo[b];
        v += prefac * (in_lo[a] * pB[b] + pB[a] * in_lo[b] - uBv * in_lo[a] * out_lo[b] -
                       uBv * out_lo[a] * in_lo[b] - vBv * in_lo[a] * in_lo[b]);
        bc_minus[a][b] = v - rhs_minus[a][b];
      }
  }
  if (in.physical) {
    // add_physical_terms_to_dt_v_minus (BjorhusImpl.cpp:223-494; mu_phys = 0,
    // adjust_phys_using_c4, gamma2_in_phys): incoming Weyl propagating mode U^{3-}
    double K[3][3], chr2[3][3][3], w[3][4], cov_dK[3][3][3], ricci[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        K[i][j] = 0.5 * in.pi[i + 1][j + 1] + 0.5 * (phi_t[i][j + 1] + phi_t[j][i + 1]);
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) {
          double v = 0.0;
          for (int j = 0; j < 3; ++j)   // Gamma_{j kl} = (d_l g_jk + d_k g_jl - d_j g_kl) / 2
            v += ig[i][j] * 0.5 * (in.phi[l][j + 1][k + 1] + in.phi[k][j + 1][l + 1] -
                                   in.phi[j][k + 1][l + 1]);
          chr2[i][k][l] = v;
        }
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < 4; ++a) {
        double v = 0.5 * in.t_up[a] * phi_tt[k];
        for (int c = 0; c < 4; ++c) v += in.ipsi[c][a] * phi_t[k][c];
        w[k][a] = v;
      }
    for (int k = 0; k < 3; ++k)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double v = in.d_pi[k][i + 1][j + 1];
          for (int a = 0; a < 4; ++a) {
            v += (in.d_phi[k][i][j + 1][a] + in.d_phi[k][j][i + 1][a]) * in.t_up[a];
            v -= (in.phi[i][j + 1][a] + in.phi[j][i + 1][a]) * w[k][a];
          }
          v *= 0.5;
          for (int l = 0; l < 3; ++l) v -= chr2[l][i][k] * K[l][j] + chr2[l][j][k] * K[l][i];
          cov_dK[k][i][j] = v;
        }
    {
      // spatial Ricci tensor of the GH variables (GeneralizedHarmonic/Ricci.cpp)
      double pI[3][3][3], pK[3][3][3], dm2b[3];   // 1/2 g^{kl} d_l g_ij, 1/2 g^{kl} d_i g_jl
      for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            double v1 = 0.0, v2 = 0.0;
            for (int l = 0; l < 3; ++l) {
              v1 += ig[k][l] * in.phi[l][i + 1][j + 1];
              v2 += ig[k][l] * in.phi[i][j + 1][l + 1];
            }
            pI[k][i][j] = 0.5 * v1;
            pK[i][j][k] = 0.5 * v2;
          }
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
        for (int l = 0; l < 3; ++l)
          for (int i = 0; i < 3; ++i) v += ig[k][l] * (pK[l][i][i] - 2.0 * pI[i][i][l]);
        dm2b[k] = v;
      }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double v = 0.0;
          for (int k = 0; k < 3; ++k)
            for (int l = 0; l < 3; ++l) {
              // d3[x][y][z][u] = d_x Phi_{y z u} (spatial)
#define D3(x, y, z, u) in.d_phi[x][y][(z) + 1][(u) + 1]
              v += 0.25 * ig[k][l] *
                   (D3(j, l, k, i) + D3(i, l, k, j) - D3(j, i, k, l) - D3(i, j, k, l) +
                    D3(k, i, j, l) + D3(k, j, i, l) - 2.0 * D3(l, k, i, j));
              // adjust_phys_using_c4
              v += 0.25 * ig[k][l] *
                   (D3(i, k, l, j) - D3(k, i, l, j) + D3(j, k, l, i) - D3(k, j, l, i));
#undef D3
              v += pK[i][k][l] * pK[j][l][k] + 2.0 * pI[k][i][l] * pK[k][j][l] -
                   2.0 * pI[k][l][i] * pI[l][k][j];
            }
          for (int k = 0; k < 3; ++k) {
            v += 0.5 * (in.phi[i][j + 1][k + 1] + in.phi[j][i + 1][k + 1] -
                        in.phi[k][i + 1][j + 1]) * dm2b[k];
            double c4t = 0.0;
            for (int a = 0; a < 4; ++a)
              c4t += in.t_up[a] * (in.d_phi[i][k][j + 1][a] - in.d_phi[k][i][j + 1][a] +
                                   in.d_phi[j][k][i + 1][a] - in.d_phi[k][j][i + 1][a]);
            v += 0.5 * n_up[k] * c4t;
          }
          ricci[i][j] = v;
        }
    }
    double trK = 0.0;
    for (int k = 0; k < 3; ++k)
      for (int l = 0; l < 3; ++l) trK += K[k][l] * ig[k][l];
    double tmp[3][3], tr_tmp = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double v = ricci[i][j] + trK * K[i][j];
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l < 3; ++l) v -= K[i][l] * ig[k][l] * K[k][j];
        // incoming mode: sign = -1 (WeylPropagating.cpp)
        for (int k = 0; k < 3; ++k)
          v += n_up[k] * (cov_dK[k][i][j] - 0.5 * cov_dK[j][i][k] - 0.5 * cov_dK[i][j][k]);
        tmp[i][j] = v;
      }
    for (int k = 0; k < 3; ++k)
      for (int l = 0; l < 3; ++l) tr_tmp += (ig[k][l] - n_up[k] * n_up[l]) * tmp[k][l];
    double weyl[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double v = 0.0;
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l < 3; ++l)
            v += ((k == i ? 1.0 : 0.0) - n_up[k] * in.n_lo[i]) *
                 ((l == j ? 1.0 : 0.0) - n_up[l] * in.n_lo[j]) * tmp[k][l];
        weyl[i][j] = v - 0.5 * tr_tmp * (in.g[i + 1][j + 1] - in.n_lo[i] * in.n_lo[j]);
      }
    double total[4][4], tr_total = 0.0;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        double u3 = 0.0, nc = 0.0;
        for (int i = 0; i < 3; ++i) {
          nc += n_up[i] * in.c3[i][a][b];
          for (int j = 0; j < 3; ++j) u3 += p_mix[i + 1][a] * p_mix[j + 1][b] * weyl[i][j];
        }
        total[a][b] = rhs_minus[a][b] + speed[3] * (2.0 * u3 - in.gamma2 * nc);
        tr_total += p_up[a][b] * total[a][b];
      }
    for (int c = 0; c < 4; ++c)
      for (int d = 0; d < 4; ++d) {
        double v = 0.0;
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) v += p_mix[a][c] * p_mix[b][d] * total[a][b];
        bc_minus[c][d] += v - 0.5 * tr_total * p_lo[c][d];
      }
  }
  // only incoming fields are corrected (Bjorhus.cpp:38-47, :345-352)
  const double k0 = speed[0] > 0.0 ? 0.0 : 1.0, k1 = speed[1] > 0.0 ? 0.0 : 1.0,
               k2 = speed[2] > 0.0 ? 0.0 : 1.0, k3 = speed[3] > 0.0 ? 0.0 : 1.0;
  // back to the evolved variables (Characteristics.cpp:133-169)
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      const double vpsi = k0 * bc_psi[a][b], vp = k2 * bc_plus[a][b], vm = k3 * bc_minus[a][b];
      out.g[a][b] = vpsi;
      out.pi[a][b] = 0.5 * (vp + vm) + in.gamma2 * vpsi;
      for (int i = 0; i < 3; ++i)
        out.phi[i][a][b] = in.n_lo[i] * 0.5 * (vp - vm) + k1 * bc_zero[i][a][b];
    }
}

}  // namespace dg
