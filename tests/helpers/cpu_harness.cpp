// Test-only host harness: runs the per-point algebra of
// spectre_b200/csrc/pointwise.cuh on the CPU so that it can be compared with
// the oracle without a GPU.  NOT part of the product (never built into
// libdgrhs.so, never imported by spectre_b200/).
#include <cmath>
using std::exp;
using std::log;
using std::sqrt;
#include "../../spectre_b200/csrc/pointwise.cuh"
#include "../../spectre_b200/csrc/bjorhus.cuh"

extern "C" {

// u [50][n], dlog [150][n] (logical derivative jhat of comp c at 3c+jhat),
// J [9][n] (jhat + 3 i), gam [3][n], H [4][n], dH [16][n] (a + 4 b)
// gauge: 0 harmonic, 1 fields (H, dH), 2 damped harmonic (dhp = {sigma_r, amp L1,
// L2, S, exp L1, L2, S}, coords [3][n])
void h_gh_volume(int n, int gauge, const double* u, const double* dlog,
                 const double* J, const double* gam, const double* H,
                 const double* dH, const double* dhp, const double* coords,
                 double* dt) {
  for (int p = 0; p < n; ++p) {
    double g[10], pi[10], phi[3][10], Jm[3][3], Q[10];
    for (int s = 0; s < 10; ++s) {
      g[s] = u[(size_t)s * n + p];
      pi[s] = u[(size_t)(10 + s) * n + p];
      for (int m = 0; m < 3; ++m) phi[m][s] = u[(size_t)(20 + m + 3 * s) * n + p];
    }
    for (int jh = 0; jh < 3; ++jh)
      for (int i = 0; i < 3; ++i) Jm[jh][i] = J[(size_t)(jh + 3 * i) * n + p];
    dg::GaugeH gh;
    for (int a = 0; a < 4; ++a) {
      gh.H[a] = H[(size_t)a * n + p];
      for (int b = 0; b < 4; ++b) gh.dH[a][b] = dH[(size_t)(a + 4 * b) * n + p];
    }
    dg::GhContext ctx;
    dg::GaugeInput gin;
    gin.fields = &gh;
    if (gauge == 2) {
      gin.dh = {dhp[0], dhp[1], dhp[2], dhp[3], (int)dhp[4], (int)dhp[5], (int)dhp[6]};
      for (int i = 0; i < 3; ++i) gin.x[i] = coords[(size_t)i * n + p];
    }
    if (gauge == 0)
      dg::gh_prologue<0>(g, pi, phi, Jm, gam[p], gam[n + p], gam[2 * n + p], gin, ctx, Q);
    else if (gauge == 1)
      dg::gh_prologue<1>(g, pi, phi, Jm, gam[p], gam[n + p], gam[2 * n + p], gin, ctx, Q);
    else
      dg::gh_prologue<2>(g, pi, phi, Jm, gam[p], gam[n + p], gam[2 * n + p], gin, ctx, Q);
    for (int s = 0; s < 10; ++s) {
      double ph[3], dgl[3], dpl[3], dphl[3][3], og, op, oph[3];
      for (int m = 0; m < 3; ++m) ph[m] = phi[m][s];
      for (int jh = 0; jh < 3; ++jh) {
        dgl[jh] = dlog[(size_t)(3 * s + jh) * n + p];
        dpl[jh] = dlog[(size_t)(3 * (10 + s) + jh) * n + p];
        for (int m = 0; m < 3; ++m)
          dphl[m][jh] = dlog[(size_t)(3 * (20 + m + 3 * s) + jh) * n + p];
      }
      dg::gh_pair_rhs(ctx, Q[s], g[s], pi[s], ph, dgl, dpl, dphl, og, op, oph);
      dt[(size_t)s * n + p] = og;
      dt[(size_t)(10 + s) * n + p] = op;
      for (int m = 0; m < 3; ++m) dt[(size_t)(20 + m + 3 * s) * n + p] = oph[m];
    }
  }
}

// two-kernel variant: context (26 values) -> streaming with inertial derivatives
void h_gh_volume_split(int n, const double* u, const double* dlog, const double* J,
                       const double* gam, double* dt) {
  for (int p = 0; p < n; ++p) {
    double g[10], pi[10], phi[3][10], Jm[3][3], Q[10], ig[6];
    for (int s = 0; s < 10; ++s) {
      g[s] = u[(size_t)s * n + p];
      pi[s] = u[(size_t)(10 + s) * n + p];
      for (int m = 0; m < 3; ++m) phi[m][s] = u[(size_t)(20 + m + 3 * s) * n + p];
    }
    for (int jh = 0; jh < 3; ++jh)
      for (int i = 0; i < 3; ++i) Jm[jh][i] = J[(size_t)(jh + 3 * i) * n + p];
    dg::GhContext ctx;
    dg::GaugeInput gin;
    gin.fields = nullptr;
    dg::gh_prologue_core<0>(g, pi, phi, gam[p], gam[n + p], gam[2 * n + p], gin, ctx, Q, ig);
    // what the streaming kernel rebuilds from g and the stored context
    dg::Geom3p1 q;
    dg::geom_from_metric(g, q);
    dg::GhStreamCtx sc;
    sc.lapse = q.lapse;
    for (int i = 0; i < 3; ++i) sc.shift[i] = q.shift[i];
    for (int i = 0; i < 6; ++i) sc.ig[i] = q.ig[i];
    sc.gamma1 = gam[n + p];
    sc.gamma2 = gam[2 * n + p];
    sc.half_pi_nn = ctx.half_pi_nn;
    for (int i = 0; i < 3; ++i) { sc.w[i] = ctx.w[i]; sc.half_phi_nn[i] = ctx.half_phi_nn[i]; }
    for (int s = 0; s < 10; ++s) {
      double ph[3], dl[5][3], di[5][3], og, op, oph[3];
      for (int m = 0; m < 3; ++m) ph[m] = phi[m][s];
      const int comps[5] = {s, 10 + s, 20 + 3 * s, 21 + 3 * s, 22 + 3 * s};
      for (int c = 0; c < 5; ++c)
        for (int jh = 0; jh < 3; ++jh) dl[c][jh] = dlog[(size_t)(3 * comps[c] + jh) * n + p];
      for (int c = 0; c < 5; ++c)
        for (int i = 0; i < 3; ++i)
          di[c][i] = Jm[0][i] * dl[c][0] + Jm[1][i] * dl[c][1] + Jm[2][i] * dl[c][2];
      double dphi[3][3];
      for (int m = 0; m < 3; ++m)
        for (int i = 0; i < 3; ++i) dphi[m][i] = di[2 + m][i];
      dg::gh_pair_rhs_inertial(sc, ctx.V, Q[s], g[s], pi[s], ph, di[0], di[1], dphi, og, op, oph);
      dt[(size_t)s * n + p] = og;
      dt[(size_t)(10 + s) * n + p] = op;
      for (int m = 0; m < 3; ++m) dt[(size_t)(20 + m + 3 * s) * n + p] = oph[m];
    }
  }
}

// faces: ui/ue [50][f]; unnorm_i/unnorm_e [3][f] (each side's own outward
// unnormalised covector); gi/ge [2][f] = gamma1, gamma2; lift_n = N
void h_gh_face(int f, int N, const double* ui, const double* ue,
               const double* unnorm_i, const double* unnorm_e, const double* gi,
               const double* ge, double* corr) {
  for (int p = 0; p < f; ++p) {
    double g_i[10], g_e[10], ni[3], ne[3];
    for (int s = 0; s < 10; ++s) {
      g_i[s] = ui[(size_t)s * f + p];
      g_e[s] = ue[(size_t)s * f + p];
    }
    for (int i = 0; i < 3; ++i) {
      ni[i] = unnorm_i[(size_t)i * f + p];
      ne[i] = unnorm_e[(size_t)i * f + p];
    }
    dg::GhFaceSide si, se;
    dg::gh_face_side(g_i, ni, gi[p], gi[f + p], si);
    dg::gh_face_side(g_e, ne, ge[p], ge[f + p], se);
    const double lift = -0.5 * (double)(N * (N - 1)) * si.mag;
    for (int s = 0; s < 10; ++s) {
      double phi_i[3], phi_e[3], cg, cp, cph[3];
      for (int m = 0; m < 3; ++m) {
        phi_i[m] = ui[(size_t)(20 + m + 3 * s) * f + p];
        phi_e[m] = ue[(size_t)(20 + m + 3 * s) * f + p];
      }
      dg::GhPairPackaged ki, ke;
      dg::gh_pair_package(si, g_i[s], ui[(size_t)(10 + s) * f + p], phi_i, ki);
      dg::gh_pair_package(se, g_e[s], ue[(size_t)(10 + s) * f + p], phi_e, ke);
      dg::gh_pair_boundary_terms(si, se, ki, ke, cg, cp, cph);
      corr[(size_t)s * f + p] = cg * lift;
      corr[(size_t)(10 + s) * f + p] = cp * lift;
      for (int m = 0; m < 3; ++m) corr[(size_t)(20 + m + 3 * s) * f + p] = cph[m] * lift;
    }
  }
}

void h_sw_volume(int n, const double* u, const double* dlog, const double* J,
                 const double* gamma2, double* dt) {
  for (int p = 0; p < n; ++p) {
    double up[5], d[5][3], Jm[3][3], out[5];
    for (int c = 0; c < 5; ++c) {
      up[c] = u[(size_t)c * n + p];
      for (int jh = 0; jh < 3; ++jh) d[c][jh] = dlog[(size_t)(3 * c + jh) * n + p];
    }
    for (int jh = 0; jh < 3; ++jh)
      for (int i = 0; i < 3; ++i) Jm[jh][i] = J[(size_t)(jh + 3 * i) * n + p];
    dg::sw_point_rhs(up, d, Jm, gamma2[p], out);
    for (int c = 0; c < 5; ++c) dt[(size_t)c * n + p] = out[c];
  }
}

void h_sw_face(int f, const double* ui, const double* ue, const double* ni,
               const double* ne, const double* g2i, const double* g2e,
               double* corr) {
  for (int p = 0; p < f; ++p) {
    double a[5], b[5], na[3], nb[3], c[5];
    for (int k = 0; k < 5; ++k) {
      a[k] = ui[(size_t)k * f + p];
      b[k] = ue[(size_t)k * f + p];
    }
    for (int i = 0; i < 3; ++i) {
      na[i] = ni[(size_t)i * f + p];
      nb[i] = ne[(size_t)i * f + p];
    }
    dg::sw_face_correction(a, g2i[p], na, b, g2e[p], nb, c);
    for (int k = 0; k < 5; ++k) corr[(size_t)k * f + p] = c[k];
  }
}

// ConstraintPreservingBjorhus at npts independent points; every array is
// [npts][...] in C order with the argument list of the reference's
// dt_*_ConstraintPreserving_static_mesh twins.
void h_bjorhus_cp(int npts, int physical, const double* n_lo, const double* g, const double* pi,
                  const double* phi, const double* x, const double* gamma1,
                  const double* gamma2, const double* lapse, const double* shift,
                  const double* ipsi, const double* t_up, const double* c3, const double* H,
                  const double* dH, const double* dt_g, const double* dt_pi,
                  const double* dt_phi, const double* d_pi, const double* d_phi,
                  double* out_g, double* out_pi, double* out_phi) {
  for (int p = 0; p < npts; ++p) {
    dg::BjorhusInput in;
    dg::BjorhusOutput out;
    auto cp = [&](double* dst, const double* src, int n) {
      for (int k = 0; k < n; ++k) dst[k] = src[(size_t)p * n + k];
    };
    cp(in.n_lo, n_lo, 3);
    cp(&in.g[0][0], g, 16);
    cp(&in.pi[0][0], pi, 16);
    cp(&in.phi[0][0][0], phi, 48);
    cp(in.x, x, 3);
    in.gamma1 = gamma1[p];
    in.gamma2 = gamma2[p];
    in.lapse = lapse[p];
    cp(in.shift, shift, 3);
    cp(&in.ipsi[0][0], ipsi, 16);
    cp(in.t_up, t_up, 4);
    cp(&in.c3[0][0][0], c3, 48);
    cp(in.H, H, 4);
    cp(&in.dH[0][0], dH, 16);
    cp(&in.dt_g[0][0], dt_g, 16);
    cp(&in.dt_pi[0][0], dt_pi, 16);
    cp(&in.dt_phi[0][0][0], dt_phi, 48);
    cp(&in.d_pi[0][0][0], d_pi, 48);
    cp(&in.d_phi[0][0][0][0], d_phi, 144);
    in.physical = physical != 0;
    dg::bjorhus_constraint_preserving(in, out);
    for (int k = 0; k < 16; ++k) {
      out_g[(size_t)p * 16 + k] = (&out.g[0][0])[k];
      out_pi[(size_t)p * 16 + k] = (&out.pi[0][0])[k];
    }
    for (int k = 0; k < 48; ++k) out_phi[(size_t)p * 48 + k] = (&out.phi[0][0][0])[k];
  }
}
}
