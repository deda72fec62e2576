// CUDA kernels of the DG right-hand side for sm_100a.
//
// Data layout in HBM (all fp64): every per-point field of the batch is a
// structure-of-arrays block  [element][component][npad]  with npad = N^3
// rounded up to a multiple of 16 doubles (128 B), so that every component row
// is 128-byte aligned: legal source for TMA bulk copies (cp.async.bulk needs
// 16 B) and for 128-bit vector accesses.  Within a row the grid index is
// i + N*(j + N*k) exactly as in the reference's Variables (xi fastest).
//
// Kernels (reference call sites: SURVEY.md 2.2 K1-K13):
//   face_kernel      K5-K8,K10,K11  both sides of every mortar are packaged in
//                    registers from the raw face values; every local interface
//                    is evaluated once and the lifted corrections of both
//                    elements go to a compact pair-major buffer
//   gh/sw_volume     K1-K3 + K12-K13 fused: TMA-staged element tiles,
//                    sum-factorised logical derivatives from shared memory,
//                    Jacobian folded into the pointwise coefficients, corrections
//                    added on face points, dt_u written once, stepper update
//                    u_new = a u + sum c_j v_j written to the second state buffer
//   lincomb_kernel   K12-K13 as a separate pass (only when fusion is disabled)
//   pack_halo        face slices for neighbours on other GPUs
//   others           stand-alone partial_derivatives, gauge sources, exponential
//                    filter, constraint norms
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pointwise.cuh"
#include "bjorhus.cuh"

namespace dg {

// --------------------------------------------------------------------------
// configuration
// --------------------------------------------------------------------------
// Measured on B200 (profiles/README.md, round 2; ms per fused launch, base -> variant):
// rows of the differentiation matrix from shared memory: N = 12 3.22 -> 3.07, N = 10 2.43 -> 2.59;
// per-warp stage release: N = 12 3.19 -> 3.14 (2.93 with both), N = 10 2.44 -> 2.38, N = 8 1.20 -> 1.21
#ifndef DG_SHARED_D_MIN_N
#define DG_SHARED_D_MIN_N 12
#endif
#ifndef DG_LATE_GAUGE_LOAD
#define DG_LATE_GAUGE_LOAD 0   // measured: more spills (312 B instead of 264 B per thread)
#endif
#ifndef DG_PARK_IN_RING
#define DG_PARK_IN_RING 1
#endif
#ifndef DG_STAGED_LOADS
#define DG_STAGED_LOADS 0   // measured: same spill count, one more L2 round trip per CTA
#endif
#ifndef DG_STREAMING_HINTS
#define DG_STREAMING_HINTS 1
#endif
#ifndef DG_Q_IN_SMEM
#define DG_Q_IN_SMEM 1
#endif
#ifndef DG_WARP_RELEASE_MIN_N
#define DG_WARP_RELEASE_MIN_N 10
#endif

template <int N>
struct Cfg {
  static constexpr int n = N * N * N;
  static constexpr int f = N * N;
  static constexpr int npad = (n + 15) / 16 * 16;
  // GH volume kernel ring: one stage = the 5 component rows of a (mu,nu) pair
  // plus that pair's block of lifted face corrections [6][5][f]
  static constexpr int stage_doubles = 5 * npad + 30 * f;
  // Volume kernels: one CTA per (element, chunk of points), one thread per
  // point, 255 registers.  Preferred: chunks of <= 128 points and TWO CTAs per
  // SM -- the two CTAs run out of phase (one in its FMA-bound prologue while the
  // other streams from shared memory), measured 1.34 -> 1.15 ms on config 2.
  // That needs two 2-stage rings per SM; larger elements (N >= 10) fall back to
  // chunks of <= 256 points and one CTA per SM.
  static constexpr int fixed128 = (10 * 128 + (N * N + 1) / 2 * 2) * 8 + 64;
  static constexpr bool two_cta = 2 * (2 * stage_doubles * 8 + fixed128 + 1024) <= 232448;
  static constexpr int chunk_max = two_cta ? 128 : 256;
  static constexpr int nchunk = (n + chunk_max - 1) / chunk_max;
  static constexpr int T = ((n + nchunk - 1) / nchunk + 31) / 32 * 32;
  static constexpr int min_blocks = two_cta ? 2 : 1;
  // N >= DG_SHARED_D_MIN_N: the rows of the differentiation matrix are read from
  // shared memory inside the pair loop (a transposed copy serves the xi direction
  // without bank conflicts) instead of living in 6N registers per thread
  static constexpr bool shared_D = N >= DG_SHARED_D_MIN_N;
  // a ring stage is released per warp (counter in shared memory; the last warp to
  // finish a pair issues the next TMA copies into its stage) instead of by a
  // CTA-wide barrier per pair
  static constexpr bool warp_release = N >= DG_WARP_RELEASE_MIN_N;
  static constexpr int fixed_bytes =
      (10 * T + (shared_D ? 2 : 1) * ((N * N + 1) / 2 * 2)) * 8 + 64;
  static constexpr int max_stages =
      (232448 / min_blocks - 1024 - fixed_bytes) / (stage_doubles * 8);
  static constexpr int nstage = max_stages >= 4 ? 4 : max_stages;
  static_assert(nstage >= 2, "element too large for the shared-memory ring");
};

__host__ __device__ constexpr int padded_points(int N) {
  return (N * N * N + 15) / 16 * 16;
}

// volume index of face point (qa, qb) of direction d = 2*dim + side
template <int N>
__device__ __forceinline__ int face_point(int d, int qa, int qb) {
  const int dim = d >> 1;
  const int fixed = (d & 1) ? N - 1 : 0;
  return dim == 0 ? fixed + N * (qa + N * qb)
                  : dim == 1 ? qa + N * (fixed + N * qb) : qa + N * (qb + N * fixed);
}

// --------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol): the face kernel lets the volume
// kernel that follows it in the stream start while its last wave drains; the
// volume kernel runs its prologue (which needs nothing the face kernel writes)
// and waits only before it fetches the face corrections.  Without the launch
// attribute both instructions do nothing.
// --------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_for_primary() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// --------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS: UBLKCP)
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src,
                                             uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
      "[%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// --------------------------------------------------------------------------
// sum-factorised logical derivatives of one component at one point from a
// shared-memory tile (K1: PartialDerivatives.tpp:316-363 without the
// transposes; the three "GEMMs" are N-term dot products along the grid lines)
// --------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void logical_derivs(const double* __restrict__ tc,
                                               int i, int j, int k,
                                               const double (&Di)[N],
                                               const double (&Dj)[N],
                                               const double (&Dk)[N],
                                               double (&d)[3]) {
  const double* row = tc + N * (j + N * k);
  const double* col = tc + i + N * N * k;
  const double* pil = tc + i + N * j;
  // N = 12: two partial sums per direction (even and odd m) halve the length of
  // the dependent DFMA chains; measured 6.61 -> 6.38 ms per fused launch on the
  // Kerr-Schild workload (4096 elements), but slower for N = 10 (3.24 -> 3.36 ms)
#ifndef DG_SPLIT_ACC_MIN_N
#define DG_SPLIT_ACC_MIN_N 12
#endif
  if constexpr (N >= DG_SPLIT_ACC_MIN_N && N % 2 == 0) {
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, o0 = 0.0, o1 = 0.0, o2 = 0.0;
    const double2* row2 = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int m = 0; m < N / 2; ++m) {
      const double2 v = row2[m];
      e0 = fma(Di[2 * m], v.x, e0);
      o0 = fma(Di[2 * m + 1], v.y, o0);
    }
#pragma unroll
    for (int m = 0; m < N / 2; ++m) {
      e1 = fma(Dj[2 * m], col[N * (2 * m)], e1);
      o1 = fma(Dj[2 * m + 1], col[N * (2 * m + 1)], o1);
    }
#pragma unroll
    for (int m = 0; m < N / 2; ++m) {
      e2 = fma(Dk[2 * m], pil[N * N * (2 * m)], e2);
      o2 = fma(Dk[2 * m + 1], pil[N * N * (2 * m + 1)], o2);
    }
    d[0] = e0 + o0;
    d[1] = e1 + o1;
    d[2] = e2 + o2;
    return;
  }
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  if constexpr (N % 2 == 0) {
    const double2* row2 = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int m = 0; m < N / 2; ++m) {
      const double2 v = row2[m];
      d0 = fma(Di[2 * m], v.x, d0);
      d0 = fma(Di[2 * m + 1], v.y, d0);
    }
  } else {
#pragma unroll
    for (int m = 0; m < N; ++m) d0 = fma(Di[m], row[m], d0);
  }
#pragma unroll
  for (int m = 0; m < N; ++m) d1 = fma(Dj[m], col[N * m], d1);
#pragma unroll
  for (int m = 0; m < N; ++m) d2 = fma(Dk[m], pil[N * N * m], d2);
  d[0] = d0;
  d[1] = d1;
  d[2] = d2;
}

// Lifted boundary corrections of the faces a point lies on are added in
// direction order xi, eta, zeta (add_slice_to_data, ApplyBoundaryCorrections.hpp:
// 1038-1043; the reference's order is a hash-map order, i.e. unspecified).
// The point's (at most three) faces are resolved once per thread:
// off[d] = offset of the point inside a [6][C][f] block for dimension d, or -1.
template <int N, int C>
struct FaceSlots {
  int off[3];
  __device__ __forceinline__ void init(int i, int j, int k) {
    constexpr int f = N * N;
    off[0] = i == 0 ? (0 * C) * f + j + N * k : (i == N - 1 ? (1 * C) * f + j + N * k : -1);
    off[1] = j == 0 ? (2 * C) * f + i + N * k : (j == N - 1 ? (3 * C) * f + i + N * k : -1);
    off[2] = k == 0 ? (4 * C) * f + i + N * j : (k == N - 1 ? (5 * C) * f + i + N * j : -1);
  }
  // N >= 2, so lower and upper face of a dimension are distinct points
  __device__ __forceinline__ double add(double v, const double* __restrict__ blk,
                                        int comp) const {
    constexpr int f = N * N;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double c = off[d] >= 0 ? blk[off[d] + comp * f] : 0.0;
      v += c;
    }
    return v;
  }
};

// --------------------------------------------------------------------------
// Stepper update fused into the volume kernels (K12+K13 without a separate
// pass): u_new = a*u + sum_j c_j v_j + c_new*dt_u, accumulated oldest term
// first exactly like lincomb_kernel, written to a second state buffer (the
// element tile of u is still being read by the other chunks of the element).
// --------------------------------------------------------------------------
struct UpdateArgs {
  double* u_new;   // nullptr: no fused update
  double a, c_new;
  int nterms;      // older terms (<= 3)
  double c[3];
  const double* v[3];
};

// the older terms are fetched early (hv) so that their latency hides behind
// the derivative work of the pair
__device__ __forceinline__ void fused_prefetch(const UpdateArgs& up, size_t idx,
                                               double (&hv)[3]) {
#pragma unroll
#if DG_STREAMING_HINTS
  // read once, never again: evict-first in L2, so that the lines that ARE reused (the element
  // tiles the seven chunk CTAs of an element stage, the spill slots) stay
  for (int j = 0; j < 3; ++j) hv[j] = (j < up.nterms) ? __ldcs(up.v[j] + idx) : 0.0;
#else
  for (int j = 0; j < 3; ++j) hv[j] = (j < up.nterms) ? __ldg(up.v[j] + idx) : 0.0;
#endif
}
__device__ __forceinline__ void fused_update(const UpdateArgs& up, size_t idx, double u,
                                             const double (&hv)[3], double dt_new) {
  double r = u * up.a;
#pragma unroll
  for (int j = 0; j < 3; ++j)
    if (j < up.nterms) r = fma(up.c[j], hv[j], r);
  r = fma(up.c_new, dt_new, r);
#if DG_STREAMING_HINTS
  __stcs(up.u_new + idx, r);
#else
  up.u_new[idx] = r;
#endif
}

// --------------------------------------------------------------------------
// GH volume kernel (K1+K2+K3+K11-add fused)
// --------------------------------------------------------------------------
struct GhVolArgs {
  const double* u;       // [E][50][npad]
  double* dt;            // [E][50][npad]
  const double* invjac;  // [E][9][npad]
  const double* stat;    // [E][3][npad]   gamma0, gamma1, gamma2
  const double* corr;    // [E][10][6][5][f] pair-major, or nullptr (volume only)
  const double* gH;      // [E][4][npad]   gauge H_a      (non-harmonic)
  const double* gdH;     // [E][16][npad]  d_a H_b, a+4b  (non-harmonic)
  const double* D;       // [N*N] row-major differentiation matrix
  const double* coords;  // [E][3][npad] (DampedHarmonic gauge)
  DampedHarmonicParams dh;
  int elem_begin;
  UpdateArgs upd;
  int prefetch_dist;  // CTAs ahead whose prologue inputs are pulled into L2 (0: off)
};

template <int N>
constexpr int gh_volume_smem_bytes() {
  return Cfg<N>::nstage * Cfg<N>::stage_doubles * 8 + Cfg<N>::fixed_bytes;
}

// Prologue of the GH volume kernels at one point: everything that needs all 50
// components (3+1 geometry, normal contractions, gauge source, Q_{mu nu}); the
// Jacobian is folded into the linear-coefficient context.  Q goes to shared
// memory (sQ_pt[s * q_stride]).
template <int N, int kGauge>
__device__ __forceinline__ void gh_point_prologue(const GhVolArgs& a, int e, int pt,
                                                  double* __restrict__ sQ_pt, int q_stride,
                                                  GhContext& ctx,
                                                  double* __restrict__ park_pt = nullptr) {
  constexpr int npad = Cfg<N>::npad;
  const double* __restrict__ ue = a.u + (size_t)e * 50 * npad;
  double g[10], pi[10], phi[3][10], Q[10], ig[6];
#if DG_STAGED_LOADS
  // The metric first: the 3+1 split is formed while only these ten values (and nothing that
  // is needed later) occupy registers; the other 40 + 20 loads are issued behind it.  Costs one
  // more L2 round trip per CTA (the inputs were prefetched into L2 by the previous wave),
  // saves the spills the single 76-load burst caused.
#pragma unroll
  for (int s = 0; s < 10; ++s) g[s] = __ldg(ue + (size_t)s * npad + pt);
  {
    Geom3p1 q0;
    geom_from_metric(g, q0);
    asm volatile("" ::"d"(q0.lapse), "d"(q0.ig[0]), "d"(q0.ig[3]), "d"(q0.ig[5]) : "memory");
  }
#pragma unroll
  for (int s = 0; s < 10; ++s) {
    pi[s] = __ldg(ue + (size_t)(10 + s) * npad + pt);
#pragma unroll
    for (int m = 0; m < 3; ++m)
      phi[m][s] = __ldg(ue + (size_t)(20 + m + 3 * s) * npad + pt);
  }
#else
#pragma unroll
  for (int s = 0; s < 10; ++s) {
    g[s] = __ldg(ue + (size_t)s * npad + pt);
    pi[s] = __ldg(ue + (size_t)(10 + s) * npad + pt);
#pragma unroll
    for (int m = 0; m < 3; ++m)
      phi[m][s] = __ldg(ue + (size_t)(20 + m + 3 * s) * npad + pt);
  }
#endif
  const double* se = a.stat + (size_t)e * 3 * npad + pt;
  const double gamma0 = __ldg(se), gamma1 = __ldg(se + npad),
               gamma2 = __ldg(se + 2 * npad);
  GaugeH gh;
  GaugeInput gin;
  gin.fields = &gh;
  if constexpr (kGauge == 2) {
    gin.dh = a.dh;
    const double* xe = a.coords + (size_t)e * 3 * npad + pt;
#pragma unroll
    for (int x = 0; x < 3; ++x) gin.x[x] = __ldg(xe + (size_t)x * npad);
  }
  if constexpr (kGauge == 1) {
    const double* he = a.gH + (size_t)e * 4 * npad + pt;
    const double* dhe = a.gdH + (size_t)e * 16 * npad + pt;
#if DG_LATE_GAUGE_LOAD
    gin.H_global = he;
    gin.dH_global = dhe;
    gin.stride = npad;
#else
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      gh.H[x] = __ldg(he + (size_t)x * npad);
#pragma unroll
      for (int y = 0; y < 4; ++y) gh.dH[x][y] = __ldg(dhe + (size_t)(x + 4 * y) * npad);
    }
#endif
  }
#if DG_Q_IN_SMEM
  (void)Q;
  if (park_pt != nullptr)
    gh_prologue_core<kGauge>(g, pi, phi, gamma0, gamma1, gamma2, gin, ctx,
                             QStrided{sQ_pt, q_stride}, ig, QStrided{park_pt, q_stride});
  else
    gh_prologue_core<kGauge>(g, pi, phi, gamma0, gamma1, gamma2, gin, ctx,
                             QStrided{sQ_pt, q_stride}, ig);
#else
  gh_prologue_core<kGauge>(g, pi, phi, gamma0, gamma1, gamma2, gin, ctx, Q, ig);
#pragma unroll
  for (int s = 0; s < 10; ++s) sQ_pt[s * q_stride] = Q[s];
#endif
  // the inverse Jacobian is fetched only now: keeping its 9 values out of
  // the register-critical part of the prologue avoids spills
  asm volatile("" ::: "memory");
  double J[3][3];
  const double* je = a.invjac + (size_t)e * 9 * npad + pt;
#pragma unroll
  for (int jh = 0; jh < 3; ++jh)
#pragma unroll
    for (int i = 0; i < 3; ++i) J[jh][i] = __ldg(je + (size_t)(jh + 3 * i) * npad);
  gh_context_set_jacobian(ctx, J, ig);
}

// kGauge: 0 Harmonic, 1 gauge fields from memory, 2 DampedHarmonic
template <int N, int kGauge>
__global__ void __launch_bounds__(Cfg<N>::T, Cfg<N>::min_blocks) gh_volume_kernel(GhVolArgs a) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad, T = Cfg<N>::T, f = N * N;
  constexpr int NS = Cfg<N>::nstage, SD = Cfg<N>::stage_doubles;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  double* sQ = ring + NS * SD;
  double* sD = sQ + 10 * T;
  constexpr bool kSharedD = Cfg<N>::shared_D;
  double* sDT = sD + (N * N + 1) / 2 * 2;  // transposed copy (kSharedD only)
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(sD + (kSharedD ? 2 : 1) * ((N * N + 1) / 2 * 2));
  unsigned int* released = reinterpret_cast<unsigned int*>(bars + 4);  // [NS] warps done

  const int e = a.elem_begin + blockIdx.x / Cfg<N>::nchunk;
  const int chunk = blockIdx.x % Cfg<N>::nchunk;
  const int tid = threadIdx.x;
  const int pt = chunk * T + tid;
  const bool active = pt < n;
  const double* __restrict__ ue = a.u + (size_t)e * 50 * npad;
  const bool with_corr = a.corr != nullptr;
  const double* __restrict__ ce = with_corr ? a.corr + (size_t)e * 10 * 30 * f : nullptr;

  // stage the 5 components of pair s = (g_s, Pi_s, Phi_0s, Phi_1s, Phi_2s)
  // and the pair's face corrections
  // (which = 1: the state tiles, 2: the corrections, 3: both; the barrier expects
  // the bytes of both from the first call on)
  auto issue = [&](int s, int stage, int which) {
    double* t = ring + stage * SD;
    if (which & 1) {
      mbar_expect_tx(&bars[stage], (5 * npad + (with_corr ? 30 * f : 0)) * 8);
      tma_bulk_g2s(t, ue + (size_t)s * npad, npad * 8, &bars[stage]);
      tma_bulk_g2s(t + npad, ue + (size_t)(10 + s) * npad, npad * 8, &bars[stage]);
      tma_bulk_g2s(t + 2 * npad, ue + (size_t)(20 + 3 * s) * npad, 3 * npad * 8,
                   &bars[stage]);
    }
    if ((which & 2) && with_corr)
      tma_bulk_g2s(t + 5 * npad, ce + (size_t)s * 30 * f, 30 * f * 8, &bars[stage]);
  };
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < NS; ++st) {
      mbar_init(&bars[st], 1);
      released[st] = 0u;
    }
    mbar_fence_init();
#pragma unroll
    for (int st = 0; st < NS; ++st) issue(st, st, 1);
  }
  for (int idx = tid; idx < N * N; idx += T) {
    sD[idx] = a.D[idx];
    if constexpr (kSharedD) sDT[(idx % N) * N + idx / N] = a.D[idx];
  }

  // ---- prologue: everything that needs all 50 components at the point ----
  // N = 12: the correction block of the second ring stage is idle until the prologue is over
  // (its copy is issued behind the barrier below): sixteen doubles per thread of it park the
  // early results of the prologue (see LocalPark) instead of registers / spill slots
  constexpr bool kPark = DG_PARK_IN_RING && NS == 2 && 30 * f >= 16 * T;
  GhContext ctx;
  if (active)
    gh_point_prologue<N, kGauge>(a, e, pt, sQ + tid, T, ctx,
                                 kPark ? ring + SD + 5 * npad + tid : nullptr);
  // the face corrections are the only input the preceding face kernel writes: under
  // programmatic dependent launch everything above overlaps that kernel's last wave
  if (tid == 0 && with_corr) {
    pdl_wait_for_primary();
#pragma unroll
    for (int st = 0; st < (kPark ? 1 : NS); ++st) issue(st, st, 2);
  }
  __syncthreads();  // sD visible, barrier init visible to all waiters
  if constexpr (kPark) {
    if (tid == 0 && with_corr) issue(1, 1, 2);
  }

  const int i = pt % N, j = (pt / N) % N, k = pt / (N * N);
  double Di[kSharedD ? 1 : N], Dj[kSharedD ? 1 : N], Dk[kSharedD ? 1 : N];
  if constexpr (!kSharedD) {
    if (active) {
#pragma unroll
      for (int m = 0; m < N; ++m) {
        Di[m] = sD[i * N + m];
        Dj[m] = sD[j * N + m];
        Dk[m] = sD[k * N + m];
      }
    }
  }
  double* __restrict__ dte = a.dt + (size_t)e * 50 * npad;
  FaceSlots<N, 5> slots;
  slots.init(i, j, k);

  // ---- stream the ten (mu,nu) pairs through the ring ----
#pragma unroll 1
  for (int s = 0; s < 10; ++s) {
    const int stage = s % NS;
    const bool do_upd = a.upd.u_new != nullptr;
    const size_t ubase = (size_t)e * 50 * npad + pt;
    double hv[5][3];
    if (active && do_upd) {
      fused_prefetch(a.upd, ubase + (size_t)s * npad, hv[0]);
      fused_prefetch(a.upd, ubase + (size_t)(10 + s) * npad, hv[1]);
#pragma unroll
      for (int m = 0; m < 3; ++m)
        fused_prefetch(a.upd, ubase + (size_t)(20 + m + 3 * s) * npad, hv[2 + m]);
    }
    if (a.prefetch_dist > 0) {
      // The CTA that will run on some SM when this one retires starts with a prologue that
      // waits on ~80 global loads per thread (12 % of the kernel's stall samples at N = 12,
      // one CTA per SM, nothing to overlap with): request its inputs into L2 now, a few
      // component rows per pair.
      const unsigned int bt = blockIdx.x + (unsigned int)a.prefetch_dist;
      if (bt < gridDim.x) {
        const int et = a.elem_begin + (int)(bt / Cfg<N>::nchunk);
        const int ptt = (int)(bt % Cfg<N>::nchunk) * T + tid;
        if (ptt < n) {
          const double* ut = a.u + (size_t)et * 50 * npad + ptt;
          prefetch_l2(ut + (size_t)s * npad);
          prefetch_l2(ut + (size_t)(10 + s) * npad);
#pragma unroll
          for (int m = 0; m < 3; ++m) prefetch_l2(ut + (size_t)(20 + m + 3 * s) * npad);
          if (s < 9)
            prefetch_l2(a.invjac + ((size_t)et * 9 + s) * npad + ptt);
          else
#pragma unroll
            for (int m = 0; m < 3; ++m) prefetch_l2(a.stat + ((size_t)et * 3 + m) * npad + ptt);
          if constexpr (kGauge == 1) {
            if (s < 4) prefetch_l2(a.gH + ((size_t)et * 4 + s) * npad + ptt);
            prefetch_l2(a.gdH + ((size_t)et * 16 + s) * npad + ptt);
            if (s < 6) prefetch_l2(a.gdH + ((size_t)et * 16 + 10 + s) * npad + ptt);
          }
          if constexpr (kGauge == 2) {
            if (s < 3) prefetch_l2(a.coords + ((size_t)et * 3 + s) * npad + ptt);
          }
        }
      }
    }
    mbar_wait(&bars[stage], (s / NS) & 1);
    const double* t = ring + stage * SD;
    if (active) {
      double dg[3], dpi[3], dph[3][3], ph[3];
      if constexpr (kSharedD) {
        // m outer, the five components inner: 15 independent FMA chains, the three
        // matrix entries of this m are fetched once for all components
        double acc[5][3];
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.0;
        const double* row = t + N * (j + N * k);
        const double* col = t + i + N * N * k;
        const double* pil = t + i + N * j;
#pragma unroll
        for (int m = 0; m < N; ++m) {
          const double di = sDT[m * N + i], dj = sD[j * N + m], dk = sD[k * N + m];
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            acc[c][0] = fma(di, row[c * npad + m], acc[c][0]);
            acc[c][1] = fma(dj, col[c * npad + N * m], acc[c][1]);
            acc[c][2] = fma(dk, pil[c * npad + N * N * m], acc[c][2]);
          }
        }
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          dg[x] = acc[0][x];
          dpi[x] = acc[1][x];
#pragma unroll
          for (int m = 0; m < 3; ++m) dph[m][x] = acc[2 + m][x];
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) ph[m] = t[(2 + m) * npad + pt];
      } else {
        logical_derivs<N>(t, i, j, k, Di, Dj, Dk, dg);
        logical_derivs<N>(t + npad, i, j, k, Di, Dj, Dk, dpi);
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          logical_derivs<N>(t + (2 + m) * npad, i, j, k, Di, Dj, Dk, dph[m]);
          ph[m] = t[(2 + m) * npad + pt];
        }
      }
      double o[5];
      {
        double oph[3];
        gh_pair_rhs(ctx, sQ[s * T + tid], t[pt], t[npad + pt], ph, dg, dpi, dph, o[0],
                    o[1], oph);
        o[2] = oph[0];
        o[3] = oph[1];
        o[4] = oph[2];
      }
      if (with_corr) {
        const double* cs = t + 5 * npad;  // [6][5][f]
#pragma unroll
        for (int c = 0; c < 5; ++c) o[c] = slots.add(o[c], cs, c);
      }
#if DG_STREAMING_HINTS
      __stcs(dte + (size_t)s * npad + pt, o[0]);
      __stcs(dte + (size_t)(10 + s) * npad + pt, o[1]);
#pragma unroll
      for (int m = 0; m < 3; ++m) __stcs(dte + (size_t)(20 + m + 3 * s) * npad + pt, o[2 + m]);
#else
      dte[(size_t)s * npad + pt] = o[0];
      dte[(size_t)(10 + s) * npad + pt] = o[1];
#pragma unroll
      for (int m = 0; m < 3; ++m) dte[(size_t)(20 + m + 3 * s) * npad + pt] = o[2 + m];
#endif
      if (do_upd) {
        fused_update(a.upd, ubase + (size_t)s * npad, t[pt], hv[0], o[0]);
        fused_update(a.upd, ubase + (size_t)(10 + s) * npad, t[npad + pt], hv[1], o[1]);
#pragma unroll
        for (int m = 0; m < 3; ++m)
          fused_update(a.upd, ubase + (size_t)(20 + m + 3 * s) * npad, ph[m], hv[2 + m],
                       o[2 + m]);
      }
    }
    if constexpr (Cfg<N>::warp_release) {
      // every lane's reads of the stage feed values it has already stored, so after
      // the warp converges the stage is free as far as this warp is concerned; the
      // last of the T/32 warps to get here refills it
      __syncwarp();
      if ((tid & 31) == 0 && s + NS < 10) {
        __threadfence_block();
        const unsigned int done = atomicAdd(&released[stage], 1u) + 1u;
        if (done == (unsigned int)((T / 32) * (s / NS + 1))) issue(s + NS, stage, 3);
      }
    } else {
      __syncthreads();  // every reader is done with this stage
      if (tid == 0 && s + NS < 10) issue(s + NS, stage, 3);
    }
  }
}

// --------------------------------------------------------------------------
// Two-kernel GH volume path (N <= 10): gh_context_kernel + gh_stream_kernel.
// The single fused kernel above needs 254 registers for its prologue and is
// therefore limited to 8 warps/SM for the streaming phase as well.  Here the
// prologue runs as its own pointwise kernel (255 registers, pure streaming)
// and leaves 26 doubles per point in HBM; the streaming kernel then fits in
// 128 registers, i.e. two CTAs (16 warps) per SM.
// --------------------------------------------------------------------------
struct GhCtxArgs {
  const double* u;
  const double* stat;
  const double* gH;
  const double* gdH;
  const double* coords;
  double* ctxbuf;  // [E][26][npad]
  DampedHarmonicParams dh;
  int elem_begin, elem_end;
};

template <int N, int kGauge>
__global__ void __launch_bounds__(256, 1) gh_context_kernel(GhCtxArgs a) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(a.elem_end - a.elem_begin) * n) return;
  const int e = a.elem_begin + (int)(idx / n);
  const int pt = (int)(idx % n);
  const double* __restrict__ ue = a.u + (size_t)e * 50 * npad + pt;
  double g[10], pi[10], phi[3][10], Q[10], ig[6];
#pragma unroll
  for (int s = 0; s < 10; ++s) {
    g[s] = __ldg(ue + (size_t)s * npad);
    pi[s] = __ldg(ue + (size_t)(10 + s) * npad);
#pragma unroll
    for (int m = 0; m < 3; ++m) phi[m][s] = __ldg(ue + (size_t)(20 + m + 3 * s) * npad);
  }
  const double* se = a.stat + (size_t)e * 3 * npad + pt;
  const double gamma0 = __ldg(se), gamma1 = __ldg(se + npad), gamma2 = __ldg(se + 2 * npad);
  GaugeH gh;
  GaugeInput gin;
  gin.fields = &gh;
  if constexpr (kGauge == 2) {
    gin.dh = a.dh;
    const double* xe = a.coords + (size_t)e * 3 * npad + pt;
#pragma unroll
    for (int x = 0; x < 3; ++x) gin.x[x] = __ldg(xe + (size_t)x * npad);
  }
  if constexpr (kGauge == 1) {
    const double* he = a.gH + (size_t)e * 4 * npad + pt;
    const double* dhe = a.gdH + (size_t)e * 16 * npad + pt;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      gh.H[x] = __ldg(he + (size_t)x * npad);
#pragma unroll
      for (int y = 0; y < 4; ++y) gh.dH[x][y] = __ldg(dhe + (size_t)(x + 4 * y) * npad);
    }
  }
  GhContext ctx;
  gh_prologue_core<kGauge>(g, pi, phi, gamma0, gamma1, gamma2, gin, ctx, Q, ig);
  double* __restrict__ out = a.ctxbuf + (size_t)e * kGhCtxComps * npad + pt;
#pragma unroll
  for (int s = 0; s < 10; ++s) out[(size_t)s * npad] = Q[s];
  out[(size_t)10 * npad] = ctx.half_pi_nn;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    out[(size_t)(11 + m) * npad] = ctx.w[m];
    out[(size_t)(14 + m) * npad] = ctx.half_phi_nn[m];
#pragma unroll
    for (int x = 0; x < 3; ++x) out[(size_t)(17 + 3 * m + x) * npad] = ctx.V[m][x];
  }
}

template <int N>
struct SCfg {
  static constexpr int T = Cfg<N>::T;
  static constexpr int npad = Cfg<N>::npad;
  static constexpr int stage_doubles = Cfg<N>::stage_doubles;
  static constexpr int nstage = 2;
  // ring | per-thread J (9) and V (9) | D | barriers
  static constexpr int smem_bytes =
      (nstage * stage_doubles + 18 * T + (N * N + 1) / 2 * 2) * 8 + 64;
  static constexpr bool fits = smem_bytes <= 232448;
  static constexpr int min_blocks = (2 * (smem_bytes + 1024) <= 233472) ? 2 : 1;
};

template <int N>
__global__ void __launch_bounds__(SCfg<N>::T, SCfg<N>::min_blocks)
    gh_stream_kernel(GhVolArgs a, const double* __restrict__ ctxbuf) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad, T = Cfg<N>::T, f = N * N;
  constexpr int NS = SCfg<N>::nstage, SD = SCfg<N>::stage_doubles;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  double* sVJ = ring + NS * SD;  // [18][T]: J (jh + 3 i), then V (3 i + m)
  double* sD = sVJ + 18 * T;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + (N * N + 1) / 2 * 2);

  const int e = a.elem_begin + blockIdx.x / Cfg<N>::nchunk;
  const int chunk = blockIdx.x % Cfg<N>::nchunk;
  const int tid = threadIdx.x;
  const int pt = chunk * T + tid;
  const bool active = pt < n;
  const double* __restrict__ ue = a.u + (size_t)e * 50 * npad;
  const bool with_corr = a.corr != nullptr;
  const double* __restrict__ ce = with_corr ? a.corr + (size_t)e * 10 * 30 * f : nullptr;

  auto issue = [&](int s, int stage) {
    double* t = ring + stage * SD;
    mbar_expect_tx(&bars[stage], (5 * npad + (with_corr ? 30 * f : 0)) * 8);
    tma_bulk_g2s(t, ue + (size_t)s * npad, npad * 8, &bars[stage]);
    tma_bulk_g2s(t + npad, ue + (size_t)(10 + s) * npad, npad * 8, &bars[stage]);
    tma_bulk_g2s(t + 2 * npad, ue + (size_t)(20 + 3 * s) * npad, 3 * npad * 8,
                 &bars[stage]);
    if (with_corr) tma_bulk_g2s(t + 5 * npad, ce + (size_t)s * 30 * f, 30 * f * 8, &bars[stage]);
  };
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < NS; ++st) mbar_init(&bars[st], 1);
    mbar_fence_init();
#pragma unroll
    for (int st = 0; st < NS; ++st) issue(st, st);
  }
  for (int idx = tid; idx < N * N; idx += T) sD[idx] = a.D[idx];

  // ---- per-point context: geometry from g, the rest from the context kernel
  GhStreamCtx c;
  const double* __restrict__ cb = ctxbuf + (size_t)e * kGhCtxComps * npad + pt;
  if (active) {
    double g[10];
#pragma unroll
    for (int s = 0; s < 10; ++s) g[s] = __ldg(ue + (size_t)s * npad + pt);
    Geom3p1 q;
    geom_from_metric(g, q);
    c.lapse = q.lapse;
#pragma unroll
    for (int x = 0; x < 3; ++x) c.shift[x] = q.shift[x];
#pragma unroll
    for (int x = 0; x < 6; ++x) c.ig[x] = q.ig[x];
    const double* se = a.stat + (size_t)e * 3 * npad + pt;
    c.gamma1 = __ldg(se + npad);
    c.gamma2 = __ldg(se + 2 * npad);
    c.half_pi_nn = __ldg(cb + (size_t)10 * npad);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      c.w[m] = __ldg(cb + (size_t)(11 + m) * npad);
      c.half_phi_nn[m] = __ldg(cb + (size_t)(14 + m) * npad);
    }
    const double* je = a.invjac + (size_t)e * 9 * npad + pt;
#pragma unroll
    for (int x = 0; x < 9; ++x) {
      sVJ[x * T + tid] = __ldg(je + (size_t)x * npad);
      sVJ[(9 + x) * T + tid] = __ldg(cb + (size_t)(17 + x) * npad);
    }
  }
  __syncthreads();  // sD and the barrier initialisation are visible

  const int i = pt % N, j = (pt / N) % N, k = pt / (N * N);
  double* __restrict__ dte = a.dt + (size_t)e * 50 * npad;
  const bool do_upd = a.upd.u_new != nullptr;
  const size_t ubase = (size_t)e * 50 * npad + pt;
  FaceSlots<N, 5> slots;
  slots.init(i, j, k);

#pragma unroll 1
  for (int s = 0; s < 10; ++s) {
    const int stage = s % NS;
    double Qs = 0.0;
    if (active) {
      Qs = __ldg(cb + (size_t)s * npad);
      if (do_upd) {
        // the older derivative terms of the fused update: warm L2 now, load later
        for (int jt = 0; jt < a.upd.nterms; ++jt) {
          prefetch_l2(a.upd.v[jt] + ubase + (size_t)s * npad);
          prefetch_l2(a.upd.v[jt] + ubase + (size_t)(10 + s) * npad);
#pragma unroll
          for (int m = 0; m < 3; ++m)
            prefetch_l2(a.upd.v[jt] + ubase + (size_t)(20 + m + 3 * s) * npad);
        }
      }
    }
    mbar_wait(&bars[stage], (s / NS) & 1);
    const double* t = ring + stage * SD;
    if (active) {
      // sum-factorised logical derivatives of the pair's five components; the
      // D rows come from shared memory and are shared by the five components
      double acc[5][3];
#pragma unroll
      for (int cc = 0; cc < 5; ++cc) acc[cc][0] = acc[cc][1] = acc[cc][2] = 0.0;
      const int row = N * (j + N * k);
      if constexpr (N % 2 == 0) {
#pragma unroll
        for (int m2 = 0; m2 < N / 2; ++m2) {
          const double2 dm = *reinterpret_cast<const double2*>(sD + i * N + 2 * m2);
#pragma unroll
          for (int cc = 0; cc < 5; ++cc) {
            const double2 v = *reinterpret_cast<const double2*>(t + cc * npad + row + 2 * m2);
            acc[cc][0] = fma(dm.x, v.x, acc[cc][0]);
            acc[cc][0] = fma(dm.y, v.y, acc[cc][0]);
          }
        }
      } else {
#pragma unroll
        for (int m = 0; m < N; ++m) {
          const double dm = sD[i * N + m];
#pragma unroll
          for (int cc = 0; cc < 5; ++cc) acc[cc][0] = fma(dm, t[cc * npad + row + m], acc[cc][0]);
        }
      }
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const double dm = sD[j * N + m];
#pragma unroll
        for (int cc = 0; cc < 5; ++cc)
          acc[cc][1] = fma(dm, t[cc * npad + i + N * (m + N * k)], acc[cc][1]);
      }
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const double dm = sD[k * N + m];
#pragma unroll
        for (int cc = 0; cc < 5; ++cc)
          acc[cc][2] = fma(dm, t[cc * npad + i + N * (j + N * m)], acc[cc][2]);
      }
      // inertial derivatives d_x = J(jhat, x) d_jhat (PartialDerivatives.tpp:79-109)
      double di[5][3];
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const double j0 = sVJ[(0 + 3 * x) * T + tid], j1 = sVJ[(1 + 3 * x) * T + tid],
                     j2 = sVJ[(2 + 3 * x) * T + tid];
#pragma unroll
        for (int cc = 0; cc < 5; ++cc) {
          double v = j0 * acc[cc][0];
          v = fma(j1, acc[cc][1], v);
          v = fma(j2, acc[cc][2], v);
          di[cc][x] = v;
        }
      }
      double V[3][3], ph[3], dphi[3][3];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        ph[m] = t[(2 + m) * npad + pt];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          V[m][x] = sVJ[(9 + 3 * m + x) * T + tid];
          dphi[m][x] = di[2 + m][x];
        }
      }
      double o[5];
      {
        double oph[3];
        gh_pair_rhs_inertial(c, V, Qs, t[pt], t[npad + pt], ph, di[0], di[1], dphi, o[0], o[1],
                             oph);
        o[2] = oph[0];
        o[3] = oph[1];
        o[4] = oph[2];
      }
      if (with_corr) {
        const double* cs = t + 5 * npad;
#pragma unroll
        for (int cc = 0; cc < 5; ++cc) o[cc] = slots.add(o[cc], cs, cc);
      }
      dte[(size_t)s * npad + pt] = o[0];
      dte[(size_t)(10 + s) * npad + pt] = o[1];
#pragma unroll
      for (int m = 0; m < 3; ++m) dte[(size_t)(20 + m + 3 * s) * npad + pt] = o[2 + m];
      if (do_upd) {
        const size_t off[5] = {(size_t)s * npad, (size_t)(10 + s) * npad,
                               (size_t)(20 + 3 * s) * npad, (size_t)(21 + 3 * s) * npad,
                               (size_t)(22 + 3 * s) * npad};
#pragma unroll
        for (int cc = 0; cc < 5; ++cc) {
          double hv[3];
          fused_prefetch(a.upd, ubase + off[cc], hv);
          fused_update(a.upd, ubase + off[cc], t[cc * npad + pt], hv, o[cc]);
        }
      }
    }
    __syncthreads();
    if (tid == 0 && s + NS < 10) issue(s + NS, stage);
  }
}

// --------------------------------------------------------------------------
// ScalarWave volume kernel (same structure, one tile of 5 components)
// --------------------------------------------------------------------------
struct SwVolArgs {
  const double* u;       // [E][5][npad]
  double* dt;            // [E][5][npad]
  const double* invjac;  // [E][9][npad]
  const double* stat;    // [E][1][npad] gamma2
  const double* corr;    // [E][6][5][f] or nullptr
  const double* D;
  int elem_begin;
  UpdateArgs upd;
};

template <int N>
constexpr int sw_volume_smem_bytes() {
  return (5 * Cfg<N>::npad + (N * N + 1) / 2 * 2) * 8 + 8;
}

template <int N>
__global__ void __launch_bounds__(Cfg<N>::T) sw_volume_kernel(SwVolArgs a) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad, T = Cfg<N>::T;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tile = reinterpret_cast<double*>(smem_raw);
  double* sD = tile + 5 * npad;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sD + (N * N + 1) / 2 * 2);
  const int e = a.elem_begin + blockIdx.x / Cfg<N>::nchunk;
  const int chunk = blockIdx.x % Cfg<N>::nchunk;
  const int tid = threadIdx.x;
  const int pt = chunk * T + tid;
  const bool active = pt < n;
  const double* __restrict__ ue = a.u + (size_t)e * 5 * npad;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, 5 * npad * 8);
    tma_bulk_g2s(tile, ue, 5 * npad * 8, bar);
  }
  for (int idx = tid; idx < N * N; idx += T) sD[idx] = a.D[idx];
  __syncthreads();
  const int i = pt % N, j = (pt / N) % N, k = pt / (N * N);
  double Di[N], Dj[N], Dk[N], J[3][3], gamma2 = 0.0;
  if (active) {
#pragma unroll
    for (int m = 0; m < N; ++m) {
      Di[m] = sD[i * N + m];
      Dj[m] = sD[j * N + m];
      Dk[m] = sD[k * N + m];
    }
    const double* je = a.invjac + (size_t)e * 9 * npad + pt;
#pragma unroll
    for (int jh = 0; jh < 3; ++jh)
#pragma unroll
      for (int x = 0; x < 3; ++x) J[jh][x] = __ldg(je + (size_t)(jh + 3 * x) * npad);
    gamma2 = __ldg(a.stat + (size_t)e * npad + pt);
  }
  double hv[5][3];
  if (active && a.upd.u_new) {
#pragma unroll
    for (int c = 0; c < 5; ++c) fused_prefetch(a.upd, ((size_t)e * 5 + c) * npad + pt, hv[c]);
  }
  mbar_wait(bar, 0);
  if (active) {
    double u[5], d[5][3], out[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      u[c] = tile[c * npad + pt];
      logical_derivs<N>(tile + c * npad, i, j, k, Di, Dj, Dk, d[c]);
    }
    sw_point_rhs(u, d, J, gamma2, out);
    double* __restrict__ dte = a.dt + (size_t)e * 5 * npad;
    const double* corr_e = a.corr ? a.corr + (size_t)e * 6 * 5 * (N * N) : nullptr;
    FaceSlots<N, 5> slots;
    slots.init(i, j, k);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      double v = out[c];
      if (corr_e) v = slots.add(v, corr_e, c);
      dte[(size_t)c * npad + pt] = v;
      if (a.upd.u_new) fused_update(a.upd, ((size_t)e * 5 + c) * npad + pt, u[c], hv[c], v);
    }
  }
}

// --------------------------------------------------------------------------
// Face kernel: one thread per (element, direction, face point)
// --------------------------------------------------------------------------
struct FaceArgs {
  const double* u;        // [E][C][npad]
  const double* invjac;   // [E][9][npad]
  const double* stat;     // [E][S][npad]
  const int32_t* nbr;     // [E][6]
  // orientation of the neighbour across each face (OrientationMap,
  // Domain/Structure/OrientationMapHelpers.cpp:25-120, restricted to the face):
  // nbr_face[e*6+d] = nd | (perm << 3), nd = the neighbour's direction that
  // touches this face, perm bit0 = swap the two face coordinates, bit1/bit2 =
  // reverse the neighbour's first/second face coordinate.  nullptr = aligned
  // (nd = d^1, perm = 0).  Tensor components are inertial, so only the point
  // index is transformed (orient_variables_on_slice).
  const int32_t* nbr_face;
  const double* ghost;    // [G][HC][f]   u (C) | J row (3) | gammas
  double* corr;           // GH: [E][10][6][5][f] pair-major; SW: [E][6][5][f]
  int nelem;
  // Interfaces between two local elements are evaluated once, by the thread of
  // the lower-side face, which writes the lifted corrections of BOTH elements
  // (the packaged data of the two sides are shared; the reference evaluates
  // dg_boundary_terms once per element and mortar, ApplyBoundaryCorrections.hpp
  // :882-886).  pass: 0 = every interface; 1 = interfaces that touch an element
  // below n_interior (+ external faces of those elements); 2 = the rest (incl.
  // ghost faces) -- used to overlap the halo exchange.
  int n_interior, pass;
  // first element of the launch grid: pass 2 only has work on elements >= n_interior
  // (both sides of its interfaces are boundary elements)
  int elem_begin;
  // DemandOutgoingCharSpeeds on external faces without a ghost state (nbr = -1),
  // GeneralizedHarmonic/BoundaryConditions/DemandOutgoingCharSpeeds.cpp:37-76:
  // violations[0] counts face points where a characteristic speed (w.r.t. the
  // outward unit normal) is negative, violations[1] holds the most negative
  // speed seen (as a double).  nullptr = no check (BoundaryCondition `None`).
  unsigned long long* violations;
  // inertial mesh velocity [E][3][npad] of a moving mesh, or nullptr (static mesh)
  const double* mesh_v;
};

// most negative value seen so far (doubles <= 0 only: their bit patterns order
// in reverse as unsigned integers)
__device__ __forceinline__ void atomic_min_negative(unsigned long long* addr, double v) {
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

// neighbour-table entry of a face that belongs to a non-conforming (2:1) mortar:
// the face kernels skip it, mortar_kernel writes its corrections
constexpr int32_t kHangingFace = INT32_MIN;
// external face with the ConstraintPreservingBjorhus boundary condition: skipped
// by the face kernels, gh_bjorhus_kernel writes its (unlifted) dt corrections
constexpr int32_t kBjorhusFace = INT32_MIN + 1;          // Type ConstraintPreserving
constexpr int32_t kBjorhusPhysicalFace = INT32_MIN + 2;  // Type ConstraintPreservingPhysical
constexpr int32_t kPMortarFace = INT32_MIN + 3;  // neighbour with a different N (pmortar_kernel)
// local time stepping: an internal face whose correction comes from the boundary histories;
// the face kernels write a zero correction (like an external face without a boundary
// condition, but without the DemandOutgoingCharSpeeds check)
constexpr int32_t kLtsHistoryFace = INT32_MIN + 4;

// neighbour-side face coordinates of our face point (qa, qb)
template <int N>
__device__ __forceinline__ void orient_face_point(int perm, int qa, int qb, int& na,
                                                  int& nb) {
  na = (perm & 1) ? qb : qa;
  nb = (perm & 1) ? qa : qb;
  if (perm & 2) na = N - 1 - na;
  if (perm & 4) nb = N - 1 - nb;
}

// returns false if this (element, direction) task is not to be evaluated;
// nd = the neighbour's direction touching this face
__device__ __forceinline__ bool face_task(const FaceArgs& a, int e, int d, int nb, int nd,
                                          bool& two_sided) {
  if (nb >= 0) {
    // a local interface is owned by its side with the smaller face id
    if (e * 6 + d > nb * 6 + nd) return false;
    const bool in_int = e < a.n_interior || nb < a.n_interior;
    if ((a.pass == 1 && !in_int) || (a.pass == 2 && in_int)) return false;
    two_sided = true;
  } else {
    const bool e_int = e < a.n_interior;
    if ((a.pass == 1 && !e_int) || (a.pass == 2 && e_int)) return false;
    two_sided = false;
  }
  return true;
}

#ifndef DG_FACE_MIN_BLOCKS
#define DG_FACE_MIN_BLOCKS 1
#endif
template <int N>
__global__ void __launch_bounds__(128, DG_FACE_MIN_BLOCKS) gh_face_kernel(FaceArgs a) {
  constexpr int npad = Cfg<N>::npad, f = N * N, HC = 55;
  pdl_launch_dependents();  // the volume kernel may start its prologue (see pdl_wait_for_primary)
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(a.nelem - a.elem_begin) * 6 * f;
  if (idx >= total) return;
  const int q = (int)(idx % f);
  const int d = (int)((idx / f) % 6);
  const int e = a.elem_begin + (int)(idx / (6 * f));
  const int qa = q % N, qb = q / N;
  const int dim = d >> 1;
  const double sign = (d & 1) ? 1.0 : -1.0;
  const int nb = a.nbr[e * 6 + d];
  if (nb == kHangingFace || nb == kBjorhusFace || nb == kBjorhusPhysicalFace ||
      nb == kPMortarFace)
    return;  // written by the mortar / Bjorhus kernel
  const int nf = a.nbr_face ? a.nbr_face[e * 6 + d] : (d ^ 1);
  const int nd = nf & 7;
  int na_, nb_;
  orient_face_point<N>(nf >> 3, qa, qb, na_, nb_);
  const int qn = na_ + N * nb_;  // the same point in the neighbour's face ordering
  bool two_sided;
  if (!face_task(a, e, d, nb, nd, two_sided)) return;
  // pair-major layout [e][pair s][direction][5][f]: the volume kernel stages
  // one pair block per TMA copy
  double* __restrict__ corr = a.corr + (size_t)e * 10 * 30 * f + (size_t)d * 5 * f + q;
  double* __restrict__ corr_nb =
      two_sided ? a.corr + (size_t)nb * 10 * 30 * f + (size_t)nd * 5 * f + qn : nullptr;
  const int dim_n = nd >> 1;
  const double sign_n = (nd & 1) ? 1.0 : -1.0;
  const int p_own = face_point<N>(d, qa, qb);
  const double* __restrict__ uo = a.u + (size_t)e * 50 * npad + p_own;
  if (nb == -1 || nb == kLtsHistoryFace) {
    // no boundary correction on this face (outflow)
#pragma unroll 1
    for (int s = 0; s < 10; ++s)
#pragma unroll
      for (int c = 0; c < 5; ++c) corr[((size_t)s * 30 + c) * f] = 0.0;
    if (a.violations && nb == -1) {
      double g[10], unn[3];
      const double* jo = a.invjac + (size_t)e * 9 * npad + p_own;
#pragma unroll
      for (int x = 0; x < 3; ++x) unn[x] = sign * __ldg(jo + (size_t)(dim + 3 * x) * npad);
#pragma unroll
      for (int s = 0; s < 10; ++s) g[s] = __ldg(uo + (size_t)s * npad);
      const double* so = a.stat + (size_t)e * 3 * npad + p_own;
      GhFaceSide sd;
      gh_face_side(g, unn, __ldg(so + npad), __ldg(so + 2 * npad), sd);
      if (a.mesh_v) {
        const double* ve = a.mesh_v + (size_t)e * 3 * npad + p_own;
        const double v[3] = {__ldg(ve), __ldg(ve + npad), __ldg(ve + 2 * npad)};
        face_side_mesh_velocity(sd, v, 1.0 + __ldg(so + npad));
      }
      double mn = fmin(fmin(sd.speed[0], sd.speed[1]), fmin(sd.speed[2], sd.speed[3]));
      if (mn < 0.0) {
        atomicAdd(a.violations, 1ULL);
        atomic_min_negative(a.violations + 1, mn);
      }
    }
    return;
  }
  const double* __restrict__ un;  // neighbour values, component stride ns
  size_t ns;
  double unn_i[3], unn_e[3], g1e, g2e;
  {
    const double* jo = a.invjac + (size_t)e * 9 * npad + p_own;
#pragma unroll
    for (int x = 0; x < 3; ++x) unn_i[x] = sign * __ldg(jo + (size_t)(dim + 3 * x) * npad);
  }
  if (nb >= 0) {
    const int p_nb = face_point<N>(nd, na_, nb_);
    un = a.u + (size_t)nb * 50 * npad + p_nb;
    ns = npad;
    const double* jn = a.invjac + (size_t)nb * 9 * npad + p_nb;
#pragma unroll
    for (int x = 0; x < 3; ++x) unn_e[x] = sign_n * __ldg(jn + (size_t)(dim_n + 3 * x) * npad);
    const double* sn = a.stat + (size_t)nb * 3 * npad + p_nb;
    g1e = __ldg(sn + npad);
    g2e = __ldg(sn + 2 * npad);
  } else {
    // ghost slot: the sender's face in ITS ordering, with its own J row
    const int gi = -(nb + 2);
    un = a.ghost + (size_t)gi * HC * f + qn;
    ns = f;
#pragma unroll
    for (int x = 0; x < 3; ++x) unn_e[x] = sign_n * __ldg(un + (size_t)(50 + x) * f);
    g1e = __ldg(un + (size_t)53 * f);
    g2e = __ldg(un + (size_t)54 * f);
  }
  const double* so = a.stat + (size_t)e * 3 * npad + p_own;
  const double g1i = __ldg(so + npad), g2i = __ldg(so + 2 * npad);
  GhFaceSide si, se;
  {
    double g[10];
#pragma unroll
    for (int s = 0; s < 10; ++s) g[s] = __ldg(uo + (size_t)s * npad);
    gh_face_side(g, unn_i, g1i, g2i, si);
#pragma unroll
    for (int s = 0; s < 10; ++s) g[s] = __ldg(un + (size_t)s * ns);
    gh_face_side(g, unn_e, g1e, g2e, se);
  }
  if (a.mesh_v) {
    // the mesh velocity is continuous across the interface: this element's value serves both
    const double* ve = a.mesh_v + (size_t)e * 3 * npad + p_own;
    const double v[3] = {__ldg(ve), __ldg(ve + npad), __ldg(ve + 2 * npad)};
    face_side_mesh_velocity(si, v, 1.0 + g1i);
    face_side_mesh_velocity(se, v, 1.0 + g1e);
  }
  const double lift = -0.5 * (double)(N * (N - 1)) * si.mag;
  const double lift_nb = -0.5 * (double)(N * (N - 1)) * se.mag;
#pragma unroll 2
  for (int s = 0; s < 10; ++s) {
    double phi_i[3], phi_e[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      phi_i[m] = __ldg(uo + (size_t)(20 + m + 3 * s) * npad);
      phi_e[m] = __ldg(un + (size_t)(20 + m + 3 * s) * ns);
    }
    GhPairPackaged ki, ke;
    gh_pair_package(si, __ldg(uo + (size_t)s * npad), __ldg(uo + (size_t)(10 + s) * npad),
                    phi_i, ki);
    gh_pair_package(se, __ldg(un + (size_t)s * ns), __ldg(un + (size_t)(10 + s) * ns),
                    phi_e, ke);
    double cg, cp, cph[3];
    gh_pair_boundary_terms(si, se, ki, ke, cg, cp, cph);
    double* cs = corr + (size_t)s * 30 * f;
    cs[0] = cg * lift;
    cs[(size_t)f] = cp * lift;
#pragma unroll
    for (int m = 0; m < 3; ++m) cs[(size_t)(2 + m) * f] = cph[m] * lift;
    if (two_sided) {
      // the neighbour's correction from the same packaged data, roles swapped
      gh_pair_boundary_terms(se, si, ke, ki, cg, cp, cph);
      double* cn = corr_nb + (size_t)s * 30 * f;
      cn[0] = cg * lift_nb;
      cn[(size_t)f] = cp * lift_nb;
#pragma unroll
      for (int m = 0; m < 3; ++m) cn[(size_t)(2 + m) * f] = cph[m] * lift_nb;
    }
  }
}

template <int N>
__global__ void __launch_bounds__(128) sw_face_kernel(FaceArgs a) {
  constexpr int npad = Cfg<N>::npad, f = N * N, HC = 9;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(a.nelem - a.elem_begin) * 6 * f;
  if (idx >= total) return;
  const int q = (int)(idx % f);
  const int d = (int)((idx / f) % 6);
  const int e = a.elem_begin + (int)(idx / (6 * f));
  const int qa = q % N, qb = q / N;
  const int dim = d >> 1;
  const double sign = (d & 1) ? 1.0 : -1.0;
  const int nb = a.nbr[e * 6 + d];
  if (nb == kHangingFace || nb == kBjorhusFace || nb == kBjorhusPhysicalFace ||
      nb == kPMortarFace)
    return;  // written by the mortar / Bjorhus kernel
  const int nf = a.nbr_face ? a.nbr_face[e * 6 + d] : (d ^ 1);
  const int nd = nf & 7;
  int na_, nb_;
  orient_face_point<N>(nf >> 3, qa, qb, na_, nb_);
  const int qn = na_ + N * nb_;
  const int dim_n = nd >> 1;
  const double sign_n = (nd & 1) ? 1.0 : -1.0;
  bool two_sided;
  if (!face_task(a, e, d, nb, nd, two_sided)) return;
  double* __restrict__ corr = a.corr + ((size_t)e * 6 + d) * 5 * f + q;
  if (nb == -1 || nb == kLtsHistoryFace) {
#pragma unroll
    for (int c = 0; c < 5; ++c) corr[(size_t)c * f] = 0.0;
    return;
  }
  const int p_own = face_point<N>(d, qa, qb);
  const double* __restrict__ uo = a.u + (size_t)e * 5 * npad + p_own;
  double ui[5], ue[5], ni[3], ne[3], g2e;
  {
    const double* jo = a.invjac + (size_t)e * 9 * npad + p_own;
#pragma unroll
    for (int x = 0; x < 3; ++x) ni[x] = sign * __ldg(jo + (size_t)(dim + 3 * x) * npad);
#pragma unroll
    for (int c = 0; c < 5; ++c) ui[c] = __ldg(uo + (size_t)c * npad);
  }
  if (nb >= 0) {
    const int p_nb = face_point<N>(nd, na_, nb_);
    const double* un = a.u + (size_t)nb * 5 * npad + p_nb;
    const double* jn = a.invjac + (size_t)nb * 9 * npad + p_nb;
#pragma unroll
    for (int c = 0; c < 5; ++c) ue[c] = __ldg(un + (size_t)c * npad);
#pragma unroll
    for (int x = 0; x < 3; ++x) ne[x] = sign_n * __ldg(jn + (size_t)(dim_n + 3 * x) * npad);
    g2e = __ldg(a.stat + (size_t)nb * npad + p_nb);
  } else {
    const int gi = -(nb + 2);
    const double* un = a.ghost + (size_t)gi * HC * f + qn;
#pragma unroll
    for (int c = 0; c < 5; ++c) ue[c] = __ldg(un + (size_t)c * f);
#pragma unroll
    for (int x = 0; x < 3; ++x) ne[x] = sign_n * __ldg(un + (size_t)(5 + x) * f);
    g2e = __ldg(un + (size_t)8 * f);
  }
  const double g2i = __ldg(a.stat + (size_t)e * npad + p_own);
  // flat-space normalisation (NormalCovectorAndMagnitude.hpp:78-90)
  double mi = sqrt(ni[0] * ni[0] + ni[1] * ni[1] + ni[2] * ni[2]);
  double me = sqrt(ne[0] * ne[0] + ne[1] * ne[1] + ne[2] * ne[2]);
  const double ii = 1.0 / mi, ie = 1.0 / me;
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    ni[x] *= ii;
    ne[x] *= ie;
  }
  double c5[5];
  double ndv_i = 0.0, ndv_e = 0.0;
  if (a.mesh_v) {
    // the mesh velocity is continuous across the interface: this element's value serves both
    const double* ve = a.mesh_v + (size_t)e * 3 * npad + p_own;
    const double v[3] = {__ldg(ve), __ldg(ve + npad), __ldg(ve + 2 * npad)};
    ndv_i = ni[0] * v[0];
    ndv_i += ni[1] * v[1];
    ndv_i += ni[2] * v[2];
    ndv_e = ne[0] * v[0];
    ndv_e += ne[1] * v[1];
    ndv_e += ne[2] * v[2];
  }
  sw_face_correction(ui, g2i, ni, ue, g2e, ne, c5, ndv_i, ndv_e);
  const double lift = -0.5 * (double)(N * (N - 1)) * mi;
#pragma unroll
  for (int c = 0; c < 5; ++c) corr[(size_t)c * f] = c5[c] * lift;
  if (two_sided) {
    double* __restrict__ corr_nb = a.corr + ((size_t)nb * 6 + nd) * 5 * f + qn;
    sw_face_correction(ue, g2e, ne, ui, g2i, ni, c5, ndv_e, ndv_i);
    const double lift_nb = -0.5 * (double)(N * (N - 1)) * me;
#pragma unroll
    for (int c = 0; c < 5; ++c) corr_nb[(size_t)c * f] = c5[c] * lift_nb;
  }
}

// --------------------------------------------------------------------------
// Local time stepping (ApplyBoundaryCorrections.hpp:797-1010 with
// local_time_stepping == true, AdamsBashforth::add_boundary_delta_impl, AdamsBashforth.cpp:
// 264-281).  The boundary history of a mortar is kept as raw face values of both elements at
// their own step times (a ring of `depth` snapshots per element; the packaged data are a
// pointwise function of them and of the static geometry):
//   lts_snapshot_kernel   the faces of the elements that were just evaluated -> ring slot
//   *_lts_boundary_kernel for the elements that finish a step: per internal face
//                         acc = sum_terms coef * lift(dg_boundary_terms(local(t_i), remote(t_j)))
//                         with the (i, j, coef) of adams_lts::lts_coefficients (host, one list
//                         per step-size level of the neighbour), terms in the reference's
//                         sorted order
//   lts_add_kernel        u += acc on the slices (add_slice_to_data), directions in order
// Not on the measured (GTS) path.
// --------------------------------------------------------------------------
struct LtsTerm {
  int lslot, rslot;  // ring slots of the local / remote snapshot
  double coef;
};
constexpr int kLtsMaxLevels = 8;

struct LtsBoundaryArgs {
  const double* fh;      // [E][depth][6][C][f]
  const double* invjac;
  const double* stat;
  const int32_t* nbr;
  const int32_t* nbr_face;
  const int32_t* level;  // [E] step-size level of every element
  const LtsTerm* terms;  // [levels][max_terms]
  int nterms[kLtsMaxLevels];
  int max_terms, depth, elem_begin, elem_end;
  double* acc;           // [E][6][C][f]
  double* u;             // state buffer of the completing elements (lts_add_kernel)
  // [E][6]: 1 = the face is in the boundary histories (mode 1: the neighbour has another
  // level; faces between elements of the same level went into the volume history, which is
  // what lts_coefficients_for_gts sums to)
  const uint8_t* in_history;
};

template <int N, int C>
__global__ void __launch_bounds__(128) lts_snapshot_kernel(const double* __restrict__ u,
                                                           double* __restrict__ fh,
                                                           const uint8_t* __restrict__ in_history,
                                                           int depth, int slot, int eb, int ee) {
  constexpr int npad = Cfg<N>::npad, f = N * N;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(ee - eb) * 6 * f) return;
  const int q = (int)(idx % f), d = (int)((idx / f) % 6), e = eb + (int)(idx / (6 * f));
  if (!in_history[e * 6 + d]) return;
  const int p = face_point<N>(d, q % N, q / N);
  const double* src = u + (size_t)e * C * npad + p;
  double* dst = fh + ((((size_t)e * depth + slot) * 6 + d) * C) * f + q;
#pragma unroll 5
  for (int c = 0; c < C; ++c) dst[(size_t)c * f] = src[(size_t)c * npad];
}

template <int N>
__global__ void __launch_bounds__(128) gh_lts_boundary_kernel(LtsBoundaryArgs a) {
  constexpr int npad = Cfg<N>::npad, f = N * N, C = 50;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(a.elem_end - a.elem_begin) * 6 * f) return;
  const int q = (int)(idx % f), d = (int)((idx / f) % 6);
  const int e = a.elem_begin + (int)(idx / (6 * f));
  const int nb = a.nbr[e * 6 + d];
  if (nb < 0 || !a.in_history[e * 6 + d]) return;  // (hanging faces: lts_mortar_kernel)
  const int qa = q % N, qb = q / N;
  const int nf = a.nbr_face ? a.nbr_face[e * 6 + d] : (d ^ 1);
  const int nd = nf & 7;
  int na_, nb_;
  orient_face_point<N>(nf >> 3, qa, qb, na_, nb_);
  const int qn = na_ + N * nb_;
  const int p_own = face_point<N>(d, qa, qb), p_nb = face_point<N>(nd, na_, nb_);
  double unn_i[3], unn_e[3];
  {
    const double sign = (d & 1) ? 1.0 : -1.0, sign_n = (nd & 1) ? 1.0 : -1.0;
    const double* jo = a.invjac + (size_t)e * 9 * npad + p_own;
    const double* jn = a.invjac + (size_t)nb * 9 * npad + p_nb;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      unn_i[x] = sign * __ldg(jo + (size_t)((d >> 1) + 3 * x) * npad);
      unn_e[x] = sign_n * __ldg(jn + (size_t)((nd >> 1) + 3 * x) * npad);
    }
  }
  const double* so = a.stat + (size_t)e * 3 * npad + p_own;
  const double* sn = a.stat + (size_t)nb * 3 * npad + p_nb;
  const double g1i = __ldg(so + npad), g2i = __ldg(so + 2 * npad);
  const double g1e = __ldg(sn + npad), g2e = __ldg(sn + 2 * npad);
  double* __restrict__ acc = a.acc + ((size_t)(e * 6 + d) * C) * f + q;
#pragma unroll 1
  for (int c = 0; c < C; ++c) acc[(size_t)c * f] = 0.0;
  const int cls = a.level[nb];
  const LtsTerm* terms = a.terms + (size_t)cls * a.max_terms;
#pragma unroll 1
  for (int t = 0; t < a.nterms[cls]; ++t) {
    const LtsTerm tm = terms[t];
    const double* ul = a.fh + ((((size_t)e * a.depth + tm.lslot) * 6 + d) * C) * f + q;
    const double* ur = a.fh + ((((size_t)nb * a.depth + tm.rslot) * 6 + nd) * C) * f + qn;
    GhFaceSide si, se;
    {
      double g[10];
#pragma unroll
      for (int s = 0; s < 10; ++s) g[s] = ul[(size_t)s * f];
      gh_face_side(g, unn_i, g1i, g2i, si);
#pragma unroll
      for (int s = 0; s < 10; ++s) g[s] = ur[(size_t)s * f];
      gh_face_side(g, unn_e, g1e, g2e, se);
    }
    const double lift = -0.5 * (double)(N * (N - 1)) * si.mag;
#pragma unroll 1
    for (int s = 0; s < 10; ++s) {
      double phi_i[3], phi_e[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        phi_i[m] = ul[(size_t)(20 + m + 3 * s) * f];
        phi_e[m] = ur[(size_t)(20 + m + 3 * s) * f];
      }
      GhPairPackaged ki, ke;
      gh_pair_package(si, ul[(size_t)s * f], ul[(size_t)(10 + s) * f], phi_i, ki);
      gh_pair_package(se, ur[(size_t)s * f], ur[(size_t)(10 + s) * f], phi_e, ke);
      double cg, cp, cph[3];
      gh_pair_boundary_terms(si, se, ki, ke, cg, cp, cph);
      acc[(size_t)s * f] += tm.coef * (cg * lift);
      acc[(size_t)(10 + s) * f] += tm.coef * (cp * lift);
#pragma unroll
      for (int m = 0; m < 3; ++m) acc[(size_t)(20 + m + 3 * s) * f] += tm.coef * (cph[m] * lift);
    }
  }
}

template <int N>
__global__ void __launch_bounds__(128) sw_lts_boundary_kernel(LtsBoundaryArgs a) {
  constexpr int npad = Cfg<N>::npad, f = N * N, C = 5;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(a.elem_end - a.elem_begin) * 6 * f) return;
  const int q = (int)(idx % f), d = (int)((idx / f) % 6);
  const int e = a.elem_begin + (int)(idx / (6 * f));
  const int nb = a.nbr[e * 6 + d];
  if (nb < 0 || !a.in_history[e * 6 + d]) return;  // (hanging faces: lts_mortar_kernel)
  const int qa = q % N, qb = q / N;
  const int nf = a.nbr_face ? a.nbr_face[e * 6 + d] : (d ^ 1);
  const int nd = nf & 7;
  int na_, nb_;
  orient_face_point<N>(nf >> 3, qa, qb, na_, nb_);
  const int qn = na_ + N * nb_;
  const int p_own = face_point<N>(d, qa, qb), p_nb = face_point<N>(nd, na_, nb_);
  double ni[3], ne[3];
  {
    const double sign = (d & 1) ? 1.0 : -1.0, sign_n = (nd & 1) ? 1.0 : -1.0;
    const double* jo = a.invjac + (size_t)e * 9 * npad + p_own;
    const double* jn = a.invjac + (size_t)nb * 9 * npad + p_nb;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      ni[x] = sign * __ldg(jo + (size_t)((d >> 1) + 3 * x) * npad);
      ne[x] = sign_n * __ldg(jn + (size_t)((nd >> 1) + 3 * x) * npad);
    }
  }
  const double g2i = __ldg(a.stat + (size_t)e * npad + p_own);
  const double g2e = __ldg(a.stat + (size_t)nb * npad + p_nb);
  const double mi = sqrt(ni[0] * ni[0] + ni[1] * ni[1] + ni[2] * ni[2]);
  const double me = sqrt(ne[0] * ne[0] + ne[1] * ne[1] + ne[2] * ne[2]);
  const double ii = 1.0 / mi, ie = 1.0 / me;
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    ni[x] *= ii;
    ne[x] *= ie;
  }
  const double lift = -0.5 * (double)(N * (N - 1)) * mi;
  double sum[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int cls = a.level[nb];
  const LtsTerm* terms = a.terms + (size_t)cls * a.max_terms;
#pragma unroll 1
  for (int t = 0; t < a.nterms[cls]; ++t) {
    const LtsTerm tm = terms[t];
    const double* ul = a.fh + ((((size_t)e * a.depth + tm.lslot) * 6 + d) * C) * f + q;
    const double* ur = a.fh + ((((size_t)nb * a.depth + tm.rslot) * 6 + nd) * C) * f + qn;
    double ui[5], ue[5], c5[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      ui[c] = ul[(size_t)c * f];
      ue[c] = ur[(size_t)c * f];
    }
    sw_face_correction(ui, g2i, ni, ue, g2e, ne, c5);
#pragma unroll
    for (int c = 0; c < 5; ++c) sum[c] += tm.coef * (c5[c] * lift);
  }
  double* __restrict__ acc = a.acc + ((size_t)(e * 6 + d) * C) * f + q;
#pragma unroll
  for (int c = 0; c < 5; ++c) acc[(size_t)c * f] = sum[c];
}

// u += acc on the slices of the internal faces, directions in ascending order at the
// points that several faces share (one thread per grid point: no race, fixed order)
template <int N>
__global__ void __launch_bounds__(256) lts_add_kernel(double* __restrict__ u,
                                                      const double* __restrict__ acc,
                                                      const uint8_t* __restrict__ in_history,
                                                      int C, int eb, int ee) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad, f = N * N;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)(ee - eb) * n) return;
  const int p = (int)(idx % n), e = eb + (int)(idx / n);
  const int i = p % N, j = (p / N) % N, k = p / (N * N);
  if (i > 0 && i < N - 1 && j > 0 && j < N - 1 && k > 0 && k < N - 1) return;
  const int ijk[3] = {i, j, k};
#pragma unroll
  for (int d = 0; d < 6; ++d) {
    const int dim = d >> 1;
    if (ijk[dim] != ((d & 1) ? N - 1 : 0)) continue;
    if (!in_history[e * 6 + d]) continue;
    const int q = dim == 0 ? j + N * k : dim == 1 ? i + N * k : i + N * j;
    const double* src = acc + ((size_t)(e * 6 + d) * C) * f + q;
    double* dst = u + (size_t)e * C * npad + p;
    for (int c = 0; c < C; ++c) dst[(size_t)c * npad] += src[(size_t)c * f];
  }
}

// --------------------------------------------------------------------------
// Non-conforming (2:1 h-refined) mortars: one CTA per coarse face, one thread
// per face / mortar point.  Reference data flow (InternalMortarDataImpl.hpp:
// 230-320, ApplyBoundaryCorrections.hpp:797-1045, MortarHelpers.hpp:74-129):
//   each side packages on its own face; the coarse side's packaged data (incl.
//   the characteristic speeds and the n_i v^+- fields) are interpolated to the
//   mortar = the fine neighbour's face (project_to_mortar: two 1-d passes with
//   the parent->child matrices); dg_boundary_terms on the mortar for both
//   elements; the fine element lifts on its face; the coarse element's
//   correction is L2-projected back (project_from_mortar: child->parent
//   matrices), lifted with |n| on the coarse face and summed over the mortars
//   of the face in the order of the mortar table (deterministic).
// ScalarWave runs through the same code as one "pair": with a flat unit normal
// and the speeds (0, 0, 1, -1) gh_pair_package is ScalarWave's dg_package_data.
// --------------------------------------------------------------------------
struct MortarArgs {
  const double* u;
  const double* invjac;
  const double* stat;
  double* corr;
  const int32_t* faces;    // [n_coarse_faces][4] = coarse element, direction, first mortar, count
  const int32_t* mortars;  // [n_mortars][4]      = fine element, direction | perm << 3, size_a, size_b
  const double* P;         // [3][N*N] parent->child, row-major [child point][parent point]
  const double* R;         // [3][N*N] child->parent, row-major [parent point][child point]
  // a side that lives on another rank: element = -(slot + 2), its face arrived in
  // ghost slot `slot` [HC][f] (u | J row | gammas, like every cut face)
  const double* ghost;
  int face_begin;          // first coarse-face group of this launch
  // local time stepping (lts.cu): mortars kept in the boundary histories are skipped here
  // (nullptr: none); elem_end > 0: only the groups whose coarse element is in the range
  const uint8_t* skip;
  int elem_begin, elem_end;
};

// dg_boundary_terms of one pair from PACKAGED values (the n_i v^+- fields are
// packaged fields of their own: on a mortar they are interpolated, not rebuilt)
// pk: 0 v_g, 1 gamma2 v_g, 2 v_plus, 3 v_minus, 4-6 v_zero, 7-9 n v_plus, 10-12 n v_minus
DG_HD void pair_boundary_terms_packaged(const double (&si)[4], const double (&ki)[13],
                                        const double (&se)[4], const double (&ke)[13],
                                        double (&c)[5]) {
  const double w_g_i = step_function(-si[0]), w_g_e = -step_function(se[0]);
  const double w_0_i = step_function(-si[1]), w_0_e = -step_function(se[1]);
  const double w_p_i = step_function(-si[2]), w_p_e = -step_function(se[2]);
  const double w_m_i = step_function(-si[3]), w_m_e = -step_function(se[3]);
  c[0] = w_g_e * ke[0] - w_g_i * ki[0];
  c[1] = 0.5 * (w_p_e * ke[2] + w_m_e * ke[3]) + w_g_e * ke[1] -
         0.5 * (w_p_i * ki[2] + w_m_i * ki[3]) - w_g_i * ki[1];
#pragma unroll
  for (int d = 0; d < 3; ++d)
    c[2 + d] = -0.5 * (w_m_e * ke[10 + d] - w_p_e * ke[7 + d]) + w_0_e * ke[4 + d] -
               0.5 * (w_p_i * ki[7 + d] - w_m_i * ki[10 + d]) - w_0_i * ki[4 + d];
}

template <int N>
constexpr int mortar_smem_bytes() {
  return (6 + 17 + 17 + 5 + 5) * N * N * 8;
}

template <int N, int kSystem>
__global__ void __launch_bounds__((N * N + 31) / 32 * 32) mortar_kernel(MortarArgs a) {
  constexpr int npad = Cfg<N>::npad, f = N * N, T = (N * N + 31) / 32 * 32;
  constexpr int C = kSystem == 1 ? 50 : 5, NP = kSystem == 1 ? 10 : 1;
  constexpr int S = kSystem == 1 ? 3 : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Mat = double[N * N];
  using Row = double[f];
  Mat* sP = reinterpret_cast<Mat*>(smem_raw);  // [3]
  Mat* sR = sP + 3;                            // [3]
  Row* sA = reinterpret_cast<Row*>(sR + 3);    // [17] coarse packaged values (+ 4 speeds)
  Row* sB = sA + 17;                           // [17] after the first interpolation pass
  Row* sE = sB + 17;                           // [5]  coarse correction on the mortar
  Row* sF = sE + 5;                            // [5]  after the first projection pass
  const int tid = threadIdx.x;
  const bool active = tid < f;
  const int qa = tid % N, qb = active ? tid / N : 0;
  const int32_t* fc = a.faces + 4 * (a.face_begin + blockIdx.x);
  const int ec = fc[0], dc = fc[1], m0 = fc[2], nm = fc[3];
  if (a.elem_end > 0 && (ec < a.elem_begin || ec >= a.elem_end)) return;
  constexpr int HC = C + 3 + (kSystem == 1 ? 2 : 1);
  for (int i = tid; i < 3 * N * N; i += T) {
    (&sP[0][0])[i] = a.P[i];
    (&sR[0][0])[i] = a.R[i];
  }

  // one side of the interface at this thread's point of face d of element e
  // (p = volume index of the point, q = its index on the face)
  auto make_side = [&](int e, int d, int p, int q, GhFaceSide& sd) {
    const double sign = (d & 1) ? 1.0 : -1.0;
    const int dim = d >> 1;
    double unn[3], g1, g2, g[10];
    if (e >= 0) {
      const double* jo = a.invjac + (size_t)e * 9 * npad + p;
#pragma unroll
      for (int x = 0; x < 3; ++x) unn[x] = sign * __ldg(jo + (size_t)(dim + 3 * x) * npad);
      const double* so = a.stat + (size_t)e * S * npad + p;
      g1 = kSystem == 1 ? __ldg(so + npad) : 0.0;
      g2 = kSystem == 1 ? __ldg(so + 2 * npad) : __ldg(so);
      if constexpr (kSystem == 1) {
        const double* uo = a.u + (size_t)e * C * npad + p;
#pragma unroll
        for (int s = 0; s < 10; ++s) g[s] = __ldg(uo + (size_t)s * npad);
      }
    } else {
      const double* gs = a.ghost + (size_t)(-(e + 2)) * HC * f + q;
#pragma unroll
      for (int x = 0; x < 3; ++x) unn[x] = sign * __ldg(gs + (size_t)(C + x) * f);
      g1 = kSystem == 1 ? __ldg(gs + (size_t)(C + 3) * f) : 0.0;
      g2 = kSystem == 1 ? __ldg(gs + (size_t)(C + 4) * f) : __ldg(gs + (size_t)(C + 3) * f);
      if constexpr (kSystem == 1) {
#pragma unroll
        for (int s = 0; s < 10; ++s) g[s] = __ldg(gs + (size_t)s * f);
      }
    }
    if constexpr (kSystem == 1) {
      gh_face_side(g, unn, g1, g2, sd);
    } else {
      // flat-space normalisation (NormalCovectorAndMagnitude.hpp:78-90), speeds
      // lambda_psi = lambda_0 = 0, lambda_+- = +-1 (UpwindPenalty.cpp:60-75)
      sd.mag = sqrt(unn[0] * unn[0] + unn[1] * unn[1] + unn[2] * unn[2]);
      const double inv = 1.0 / sd.mag;
#pragma unroll
      for (int x = 0; x < 3; ++x) sd.n_lo[x] = sd.n_up[x] = unn[x] * inv;
      sd.gamma2 = g2;
      sd.speed[0] = 0.0;
      sd.speed[1] = 0.0;
      sd.speed[2] = 1.0;
      sd.speed[3] = -1.0;
    }
  };
  // packaged values of pair s on one side
  auto package = [&](const GhFaceSide& sd, int e, int p, int q, int s, double (&pk)[13]) {
    const double* uo = e >= 0 ? a.u + (size_t)e * C * npad + p
                              : a.ghost + (size_t)(-(e + 2)) * HC * f + q;
    const size_t cs = e >= 0 ? (size_t)npad : (size_t)f;   // component stride
    double g, pi, ph[3];
    if constexpr (kSystem == 1) {
      g = __ldg(uo + (size_t)s * cs);
      pi = __ldg(uo + (size_t)(10 + s) * cs);
#pragma unroll
      for (int m = 0; m < 3; ++m) ph[m] = __ldg(uo + (size_t)(20 + m + 3 * s) * cs);
    } else {
      g = __ldg(uo);
      pi = __ldg(uo + cs);
#pragma unroll
      for (int m = 0; m < 3; ++m) ph[m] = __ldg(uo + (size_t)(2 + m) * cs);
    }
    GhPairPackaged k;
    gh_pair_package(sd, g, pi, ph, k);
    pk[0] = k.v_g;
    pk[1] = k.g2_v_g;
    pk[2] = k.v_plus;
    pk[3] = k.v_minus;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      pk[4 + m] = k.v_zero[m];
      pk[7 + m] = k.v_plus * sd.n_lo[m];
      pk[10 + m] = k.v_minus * sd.n_lo[m];
    }
  };
  auto corr_ptr = [&](int e, int s, int d) {
    return kSystem == 1 ? a.corr + (size_t)e * 10 * 30 * f + (size_t)s * 30 * f + (size_t)d * 5 * f
                        : a.corr + ((size_t)e * 6 + d) * 5 * f;
  };

  GhFaceSide sC;
  const int pC = active ? face_point<N>(dc, qa, qb) : 0;
  if (active) {
    make_side(ec, dc, pC, tid, sC);
#pragma unroll
    for (int x = 0; x < 4; ++x) sA[13 + x][tid] = sC.speed[x];
  }
  const double liftC = active ? -0.5 * (double)(N * (N - 1)) * sC.mag : 0.0;
  // the coarse face's correction is the sum over its mortars: start from zero
  if (active && ec >= 0) {
#pragma unroll 1
    for (int s = 0; s < NP; ++s) {
      double* cc = corr_ptr(ec, s, dc) + tid;
#pragma unroll
      for (int c = 0; c < 5; ++c) cc[(size_t)c * f] = 0.0;
    }
  }
  __syncthreads();

#pragma unroll 1
  for (int mi = 0; mi < nm; ++mi) {
    if (a.skip && a.skip[m0 + mi]) continue;
    const int32_t* mt = a.mortars + 4 * (m0 + mi);
    // mt[1] = fine direction | perm << 3: perm takes this thread's mortar point, given
    // in the coarse element's face frame, to the fine element's face point (blocks that
    // are not aligned; orient_variables_on_slice of the exchanged mortar data)
    const int ef = mt[0], df = mt[1] & 7, sa = mt[2], sb = mt[3];
    int fa = qa, fb = qb;
    if (mt[1] >> 3) orient_face_point<N>(mt[1] >> 3, qa, qb, fa, fb);
    const int qF = active ? fa + N * fb : 0;
    const double* Pa = sP[sa];
    const double* Pb = sP[sb];
    const double* Ra = sR[sa];
    const double* Rb = sR[sb];
    GhFaceSide sFn;
    const int pF = active ? face_point<N>(df, fa, fb) : 0;
    if (active) make_side(ef, df, pF, qF, sFn);
    const double liftF = active ? -0.5 * (double)(N * (N - 1)) * sFn.mag : 0.0;
    double spC[4] = {0.0, 0.0, 0.0, 0.0};  // coarse characteristic speeds on the mortar
#pragma unroll 1
    for (int s = 0; s < NP; ++s) {
      const int c_lo = 0, c_hi = s == 0 ? 17 : 13;
      if (active) {
        double pk[13];
        package(sC, ec, pC, tid, s, pk);
#pragma unroll
        for (int c = 0; c < 13; ++c) sA[c][tid] = pk[c];
      }
      __syncthreads();
      // project_to_mortar, first face dimension: thread = (a', b)
      if (active) {
        for (int c = c_lo; c < c_hi; ++c) {
          double v = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) v += Pa[qa * N + m] * sA[c][m + N * qb];
          sB[c][tid] = v;
        }
      }
      __syncthreads();
      double cF[5], cC[5];
      if (active) {
        // second face dimension: thread = (a', b')
        double pkC[13];
#pragma unroll
        for (int c = 0; c < 13; ++c) {
          double v = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) v += Pb[qb * N + m] * sB[c][qa + N * m];
          pkC[c] = v;
        }
        if (s == 0) {
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < N; ++m) v += Pb[qb * N + m] * sB[13 + x][qa + N * m];
            spC[x] = v;
          }
        }
        double pkF[13];
        package(sFn, ef, pF, qF, s, pkF);
        pair_boundary_terms_packaged(sFn.speed, pkF, spC, pkC, cF);
        pair_boundary_terms_packaged(spC, pkC, sFn.speed, pkF, cC);
        double* cf = corr_ptr(ef >= 0 ? ef : 0, s, df) + qF;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          if (ef >= 0) cf[(size_t)c * f] = cF[c] * liftF;   // (a remote fine side: its rank does it)
          sE[c][tid] = cC[c];
        }
      }
      __syncthreads();
      if (ec < 0) continue;   // remote coarse side: its rank projects and lifts (uniform branch)
      // project_from_mortar, first face dimension: thread = (a, b')
      if (active) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          double v = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) v += Ra[qa * N + m] * sE[c][m + N * qb];
          sF[c][tid] = v;
        }
      }
      __syncthreads();
      if (active) {
        double* cc = corr_ptr(ec, s, dc) + tid;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          double v = 0.0;
#pragma unroll
          for (int m = 0; m < N; ++m) v += Rb[qb * N + m] * sF[c][qa + N * m];
          v *= liftC;
          cc[(size_t)c * f] += v;
        }
      }
    }
  }
}

// --------------------------------------------------------------------------
// Local time stepping across non-conforming (2:1) mortars: mortar_kernel's data flow with
// both sides read from the snapshot rings at the slots of a coefficient list, for the side
// that completes a step (role 0: the coarse elements of [elem_begin, elem_end), role 1: the
// fine ones).  Per mortar and term, in the reference's order: package on each face,
// project_to_mortar, dg_boundary_terms, (coarse side: project_from_mortar,) lift, times the
// coefficient, summed (AdamsBashforth.cpp:264-281 on the mortar's BoundaryHistory; a coarse
// face sums over its mortars as the reference adds one mortar's delta after the other).
// --------------------------------------------------------------------------
struct LtsMortarArgs {
  const double* fh;        // [E][depth][6][C][f]
  const double* invjac;
  const double* stat;
  const int32_t* faces;    // [groups][4] as MortarArgs
  const int32_t* mortars;  // [n_mortars][4]
  const uint8_t* mortar_in_history;  // [n_mortars]
  const double* P;
  const double* R;
  const int32_t* level;
  const LtsTerm* terms;    // [levels][max_terms]: local = the completing side
  int nterms[kLtsMaxLevels];
  int max_terms, depth, elem_begin, elem_end, role;
  double* acc;             // [E][6][C][f]
};

template <int N, int kSystem>
__global__ void __launch_bounds__((N * N + 31) / 32 * 32) lts_mortar_kernel(LtsMortarArgs a) {
  constexpr int npad = Cfg<N>::npad, f = N * N, T = (N * N + 31) / 32 * 32;
  constexpr int C = kSystem == 1 ? 50 : 5, NP = kSystem == 1 ? 10 : 1;
  constexpr int S = kSystem == 1 ? 3 : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Mat = double[N * N];
  using Row = double[f];
  Mat* sP = reinterpret_cast<Mat*>(smem_raw);  // [3]
  Mat* sR = sP + 3;                            // [3]
  Row* sA = reinterpret_cast<Row*>(sR + 3);    // [17]
  Row* sB = sA + 17;                           // [17]
  Row* sE = sB + 17;                           // [5]
  Row* sF = sE + 5;                            // [5]
  const int tid = threadIdx.x;
  const bool active = tid < f;
  const int qa = tid % N, qb = active ? tid / N : 0;
  const int32_t* fc = a.faces + 4 * blockIdx.x;
  const int ec = fc[0], dc = fc[1], m0 = fc[2], nm = fc[3];
  const bool coarse_completes = a.role == 0;
  if (coarse_completes && (ec < a.elem_begin || ec >= a.elem_end)) return;
  for (int i = tid; i < 3 * N * N; i += T) {
    (&sP[0][0])[i] = a.P[i];
    (&sR[0][0])[i] = a.R[i];
  }
  auto snapshot = [&](int e, int d, int slot) {
    return a.fh + ((((size_t)e * a.depth + slot) * 6 + d) * C) * f;
  };
  // one side at this thread's point (p volume index for the static geometry, q face index)
  auto make_side = [&](int e, int d, int p, int q, const double* face, GhFaceSide& sd) {
    const double sign = (d & 1) ? 1.0 : -1.0;
    const int dim = d >> 1;
    double unn[3];
    const double* jo = a.invjac + (size_t)e * 9 * npad + p;
#pragma unroll
    for (int x = 0; x < 3; ++x) unn[x] = sign * __ldg(jo + (size_t)(dim + 3 * x) * npad);
    const double* so = a.stat + (size_t)e * S * npad + p;
    if constexpr (kSystem == 1) {
      double g[10];
#pragma unroll
      for (int s = 0; s < 10; ++s) g[s] = face[(size_t)s * f + q];
      gh_face_side(g, unn, __ldg(so + npad), __ldg(so + 2 * npad), sd);
    } else {
      sd.mag = sqrt(unn[0] * unn[0] + unn[1] * unn[1] + unn[2] * unn[2]);
      const double inv = 1.0 / sd.mag;
#pragma unroll
      for (int x = 0; x < 3; ++x) sd.n_lo[x] = sd.n_up[x] = unn[x] * inv;
      sd.gamma2 = __ldg(so);
      sd.speed[0] = 0.0;
      sd.speed[1] = 0.0;
      sd.speed[2] = 1.0;
      sd.speed[3] = -1.0;
    }
  };
  auto package = [&](const GhFaceSide& sd, const double* face, int q, int s, double (&pk)[13]) {
    double g, pi, ph[3];
    if constexpr (kSystem == 1) {
      g = face[(size_t)s * f + q];
      pi = face[(size_t)(10 + s) * f + q];
#pragma unroll
      for (int m = 0; m < 3; ++m) ph[m] = face[(size_t)(20 + m + 3 * s) * f + q];
    } else {
      g = face[q];
      pi = face[(size_t)f + q];
#pragma unroll
      for (int m = 0; m < 3; ++m) ph[m] = face[(size_t)(2 + m) * f + q];
    }
    GhPairPackaged k;
    gh_pair_package(sd, g, pi, ph, k);
    pk[0] = k.v_g;
    pk[1] = k.g2_v_g;
    pk[2] = k.v_plus;
    pk[3] = k.v_minus;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      pk[4 + m] = k.v_zero[m];
      pk[7 + m] = k.v_plus * sd.n_lo[m];
      pk[10 + m] = k.v_minus * sd.n_lo[m];
    }
  };
  // component c of pair s in the [E][6][C][f] accumulator
  auto comp = [&](int s, int c) {
    if constexpr (kSystem == 1)
      return c == 0 ? s : c == 1 ? 10 + s : 20 + (c - 2) + 3 * s;
    else
      return c;
  };
  const int pC = active ? face_point<N>(dc, qa, qb) : 0;
  double* accC = a.acc + ((size_t)(ec * 6 + dc) * C) * f + tid;
  if (coarse_completes && active)
    for (int c = 0; c < C; ++c) accC[(size_t)c * f] = 0.0;
  __syncthreads();

#pragma unroll 1
  for (int mi = 0; mi < nm; ++mi) {
    if (!a.mortar_in_history[m0 + mi]) continue;
    const int32_t* mt = a.mortars + 4 * (m0 + mi);
    const int ef = mt[0], df = mt[1] & 7, sa = mt[2], sb = mt[3];
    if (!coarse_completes && (ef < a.elem_begin || ef >= a.elem_end)) continue;
    int fa = qa, fb = qb;
    if (mt[1] >> 3) orient_face_point<N>(mt[1] >> 3, qa, qb, fa, fb);
    const int qF = active ? fa + N * fb : 0;
    const int pF = active ? face_point<N>(df, fa, fb) : 0;
    const double* Pa = sP[sa];
    const double* Pb = sP[sb];
    const double* Ra = sR[sa];
    const double* Rb = sR[sb];
    double* accF = a.acc + ((size_t)(ef * 6 + df) * C) * f + qF;
    if (!coarse_completes && active)
      for (int c = 0; c < C; ++c) accF[(size_t)c * f] = 0.0;
    // the coefficient list of the completing side against the other side's level
    const int cls = a.level[coarse_completes ? ef : ec];
    const LtsTerm* terms = a.terms + (size_t)cls * a.max_terms;
#pragma unroll 1
    for (int t = 0; t < a.nterms[cls]; ++t) {
      const LtsTerm tm = terms[t];
      const double* faceC = snapshot(ec, dc, coarse_completes ? tm.lslot : tm.rslot);
      const double* faceF = snapshot(ef, df, coarse_completes ? tm.rslot : tm.lslot);
      GhFaceSide sC, sFn;
      if (active) {
        make_side(ec, dc, pC, tid, faceC, sC);
        make_side(ef, df, pF, qF, faceF, sFn);
      }
      __syncthreads();   // the previous term is done with sA
      if (active) {
#pragma unroll
        for (int x = 0; x < 4; ++x) sA[13 + x][tid] = sC.speed[x];
      }
      const double liftC = active ? -0.5 * (double)(N * (N - 1)) * sC.mag : 0.0;
      const double liftF = active ? -0.5 * (double)(N * (N - 1)) * sFn.mag : 0.0;
      double spC[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
      for (int s = 0; s < NP; ++s) {
        const int c_hi = s == 0 ? 17 : 13;
        if (active) {
          double pk[13];
          package(sC, faceC, tid, s, pk);
#pragma unroll
          for (int c = 0; c < 13; ++c) sA[c][tid] = pk[c];
        }
        __syncthreads();
        if (active) {
          for (int c = 0; c < c_hi; ++c) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < N; ++m) v += Pa[qa * N + m] * sA[c][m + N * qb];
            sB[c][tid] = v;
          }
        }
        __syncthreads();
        if (active) {
          double pkC[13];
#pragma unroll
          for (int c = 0; c < 13; ++c) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < N; ++m) v += Pb[qb * N + m] * sB[c][qa + N * m];
            pkC[c] = v;
          }
          if (s == 0) {
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              double v = 0.0;
#pragma unroll
              for (int m = 0; m < N; ++m) v += Pb[qb * N + m] * sB[13 + x][qa + N * m];
              spC[x] = v;
            }
          }
          double pkF[13], cc[5];
          package(sFn, faceF, qF, s, pkF);
          if (coarse_completes) {
            pair_boundary_terms_packaged(spC, pkC, sFn.speed, pkF, cc);
#pragma unroll
            for (int c = 0; c < 5; ++c) sE[c][tid] = cc[c];
          } else {
            pair_boundary_terms_packaged(sFn.speed, pkF, spC, pkC, cc);
#pragma unroll
            for (int c = 0; c < 5; ++c) accF[(size_t)comp(s, c) * f] += tm.coef * (cc[c] * liftF);
          }
        }
        __syncthreads();
        if (!coarse_completes) continue;
        if (active) {
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < N; ++m) v += Ra[qa * N + m] * sE[c][m + N * qb];
            sF[c][tid] = v;
          }
        }
        __syncthreads();
        if (active) {
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < N; ++m) v += Rb[qb * N + m] * sF[c][qa + N * m];
            accC[(size_t)comp(s, c) * f] += tm.coef * (v * liftC);
          }
        }
      }
    }
  }
}

// --------------------------------------------------------------------------
// p-nonconforming faces: the neighbour has a different number of grid
// points NB per dimension and lives in another context (one context per N; its face arrives
// as the usual 55-/9-component halo on ITS NB x NB face points).  The mortar mesh has the
// larger extents NM = max(N, NB) (dg::mortar_mesh, MortarHelpers.cpp:22-49).  Both sides
// package on their own face mesh and project to the mortar (project_to_mortar: the
// interpolation of Projection.cpp:279-362, identity for the side that already has NM
// points); the boundary correction is evaluated on the mortar points, projected back to
// this element's face mesh (project_from_mortar: the L2 projection of Projection.cpp:57-262,
// identity if N == NM), lifted with this face's normal magnitude and written to the face's
// correction slots (ApplyBoundaryCorrections.hpp:286-380).  One CTA per face.
// --------------------------------------------------------------------------
struct PMortarArgs {
  const double* u;
  const double* invjac;
  const double* stat;
  double* corr;
  const int32_t* faces;  // [n][4] = element, direction, NB, neighbour direction | perm << 3
  const double* ghost;   // [n][HC][144]: the neighbour's face on its own NB x NB points
  const double* P;       // [13][144]: row NB = interpolation min(N, NB) -> NM points, [NM][lo]
  const double* R;       // [13][144]: row NB = projection NM -> min(N, NB) points, [lo][NM]
};

constexpr int kPMortarThreads = 160;
constexpr int pmortar_smem_bytes() { return (4 * 17 + 2 * 5 + 2) * 144 * 8; }

template <int N, int kSystem>
__global__ void __launch_bounds__(kPMortarThreads) pmortar_kernel(PMortarArgs a) {
  constexpr int npad = Cfg<N>::npad, f = N * N, T = kPMortarThreads;
  constexpr int C = kSystem == 1 ? 50 : 5, NP = kSystem == 1 ? 10 : 1;
  constexpr int S = kSystem == 1 ? 3 : 1;
  constexpr int HC = C + 3 + (kSystem == 1 ? 2 : 1);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Row = double[144];
  Row* sO = reinterpret_cast<Row*>(smem_raw);  // [17] own packaged values (+ 4 speeds), own face points
  Row* sN = sO + 17;                           // [17] the neighbour's, on its face points
  Row* sB = sN + 17;                           // [17] after the first interpolation pass
  Row* sM = sB + 17;                           // [17] the interpolated side on the mortar
  Row* sE = sM + 17;                           // [5]  this element's correction on the mortar
  Row* sF = sE + 5;                            // [5]  after the first projection pass
  double* sP = &sF[5][0];
  double* sR = sP + 144;
  const int tid = threadIdx.x;
  const int32_t* fc = a.faces + 4 * blockIdx.x;
  const int e = fc[0], d = fc[1], NB = fc[2], dn = fc[3] & 7, perm = fc[3] >> 3;
  const int NM = NB > N ? NB : N, fB = NB * NB, fM = NM * NM;
  const int lo = NB > N ? N : NB;   // the side that is interpolated has lo points
  for (int i = tid; i < 144; i += T) {
    sP[i] = a.P[NB * 144 + i];
    sR[i] = a.R[NB * 144 + i];
  }
  const double* gs = a.ghost + (size_t)blockIdx.x * HC * 144;

  // ---- the two sides at their own face points -------------------------------------
  GhFaceSide sdO, sdN;
  const bool own_pt = tid < f, nbr_pt = tid < fB;
  const int pO = own_pt ? face_point<N>(d, tid % N, tid / N) : 0;
  auto finish_side = [&](const double (&unn)[3], double g1, double g2, const double (&g)[10],
                         GhFaceSide& sd) {
    if constexpr (kSystem == 1) {
      gh_face_side(g, unn, g1, g2, sd);
    } else {
      sd.mag = sqrt(unn[0] * unn[0] + unn[1] * unn[1] + unn[2] * unn[2]);
      const double inv = 1.0 / sd.mag;
#pragma unroll
      for (int x = 0; x < 3; ++x) sd.n_lo[x] = sd.n_up[x] = unn[x] * inv;
      sd.gamma2 = g2;
      sd.speed[0] = 0.0;
      sd.speed[1] = 0.0;
      sd.speed[2] = 1.0;
      sd.speed[3] = -1.0;
    }
  };
  if (own_pt) {
    const double sign = (d & 1) ? 1.0 : -1.0;
    double unn[3], g[10] = {};
    const double* jo = a.invjac + (size_t)e * 9 * npad + pO;
#pragma unroll
    for (int x = 0; x < 3; ++x) unn[x] = sign * __ldg(jo + (size_t)((d >> 1) + 3 * x) * npad);
    const double* so = a.stat + (size_t)e * S * npad + pO;
    const double g1 = kSystem == 1 ? __ldg(so + npad) : 0.0;
    const double g2 = kSystem == 1 ? __ldg(so + 2 * npad) : __ldg(so);
    if constexpr (kSystem == 1) {
      const double* uo = a.u + (size_t)e * C * npad + pO;
#pragma unroll
      for (int s = 0; s < 10; ++s) g[s] = __ldg(uo + (size_t)s * npad);
    }
    finish_side(unn, g1, g2, g, sdO);
#pragma unroll
    for (int x = 0; x < 4; ++x) sO[13 + x][tid] = sdO.speed[x];
  }
  if (nbr_pt) {
    const double sign = (dn & 1) ? 1.0 : -1.0;
    double unn[3], g[10] = {};
#pragma unroll
    for (int x = 0; x < 3; ++x) unn[x] = sign * __ldg(gs + (size_t)(C + x) * fB + tid);
    const double g1 = kSystem == 1 ? __ldg(gs + (size_t)(C + 3) * fB + tid) : 0.0;
    const double g2 = kSystem == 1 ? __ldg(gs + (size_t)(C + 4) * fB + tid)
                                   : __ldg(gs + (size_t)(C + 3) * fB + tid);
    if constexpr (kSystem == 1) {
#pragma unroll
      for (int s = 0; s < 10; ++s) g[s] = __ldg(gs + (size_t)s * fB + tid);
    }
    finish_side(unn, g1, g2, g, sdN);
#pragma unroll
    for (int x = 0; x < 4; ++x) sN[13 + x][tid] = sdN.speed[x];
  }
  const double liftO = own_pt ? -0.5 * (double)(N * (N - 1)) * sdO.mag : 0.0;
  auto package = [&](const GhFaceSide& sd, const double* base, size_t cs, int s, Row* dst) {
    double g, pi, ph[3];
    if constexpr (kSystem == 1) {
      g = __ldg(base + (size_t)s * cs);
      pi = __ldg(base + (size_t)(10 + s) * cs);
#pragma unroll
      for (int m = 0; m < 3; ++m) ph[m] = __ldg(base + (size_t)(20 + m + 3 * s) * cs);
    } else {
      g = __ldg(base);
      pi = __ldg(base + cs);
#pragma unroll
      for (int m = 0; m < 3; ++m) ph[m] = __ldg(base + (size_t)(2 + m) * cs);
    }
    GhPairPackaged k;
    gh_pair_package(sd, g, pi, ph, k);
    dst[0][tid] = k.v_g;
    dst[1][tid] = k.g2_v_g;
    dst[2][tid] = k.v_plus;
    dst[3][tid] = k.v_minus;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      dst[4 + m][tid] = k.v_zero[m];
      dst[7 + m][tid] = k.v_plus * sd.n_lo[m];
      dst[10 + m][tid] = k.v_minus * sd.n_lo[m];
    }
  };
  // this thread's mortar point (a, b) in this element's face frame and the same point in
  // the neighbour's frame (orient_variables_on_slice of the received data)
  const bool mortar_pt = tid < fM;
  const int ma = tid % NM, mb = mortar_pt ? tid / NM : 0;
  int na = (perm & 1) ? mb : ma, nb = (perm & 1) ? ma : mb;
  if (perm & 2) na = NM - 1 - na;
  if (perm & 4) nb = NM - 1 - nb;
  const int qN = na + NM * nb;
  const bool own_is_low = N < NB;
  double spM[4] = {0.0, 0.0, 0.0, 0.0};  // speeds of the interpolated side on the mortar
  __syncthreads();

#pragma unroll 1
  for (int s = 0; s < NP; ++s) {
    const int c_hi = s == 0 ? 17 : 13;
    if (own_pt) package(sdO, a.u + (size_t)e * C * npad + pO, (size_t)npad, s, sO);
    if (nbr_pt) package(sdN, gs + tid, (size_t)fB, s, sN);
    __syncthreads();
    // project_to_mortar of the side with fewer points, in that side's own frame:
    // first face dimension (thread = (a', b), a' < NM, b < lo), then the second
    const Row* src = own_is_low ? sO : sN;
    if (tid < NM * lo) {
      const int a2 = tid % NM, b = tid / NM;
      for (int c = 0; c < c_hi; ++c) {
        double v = 0.0;
        for (int m = 0; m < lo; ++m) v += sP[a2 * lo + m] * src[c][m + lo * b];
        sB[c][tid] = v;
      }
    }
    __syncthreads();
    if (mortar_pt) {
      const int a2 = tid % NM, b2 = tid / NM;
      for (int c = 0; c < 13; ++c) {
        double v = 0.0;
        for (int m = 0; m < lo; ++m) v += sP[b2 * lo + m] * sB[c][a2 + NM * m];
        sM[c][tid] = v;
      }
      if (s == 0)
        for (int x = 0; x < 4; ++x) {
          double v = 0.0;
          for (int m = 0; m < lo; ++m) v += sP[b2 * lo + m] * sB[13 + x][a2 + NM * m];
          sM[13 + x][tid] = v;
        }
    }
    __syncthreads();
    if (mortar_pt) {
      double pkO[13], pkN[13], spO[4], spN[4], cO[5];
      const Row* own = own_is_low ? sM : sO;   // on the mortar, this element's frame
      const Row* nbr = own_is_low ? sN : sM;   // on the mortar, the neighbour's frame
#pragma unroll
      for (int c = 0; c < 13; ++c) {
        pkO[c] = own[c][tid];
        pkN[c] = nbr[c][qN];
      }
      if (s == 0) {
#pragma unroll
        for (int x = 0; x < 4; ++x) spM[x] = sM[13 + x][own_is_low ? tid : qN];
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        spO[x] = own_is_low ? spM[x] : sO[13 + x][tid];
        spN[x] = own_is_low ? sN[13 + x][qN] : spM[x];
      }
      pair_boundary_terms_packaged(spO, pkO, spN, pkN, cO);
#pragma unroll
      for (int c = 0; c < 5; ++c) sE[c][tid] = cO[c];
    }
    __syncthreads();
    double* cc = (kSystem == 1 ? a.corr + (size_t)e * 10 * 30 * f + (size_t)s * 30 * f + (size_t)d * 5 * f
                               : a.corr + ((size_t)e * 6 + d) * 5 * f) + tid;
    if (!own_is_low) {
      // this face is the mortar: lift (LiftFlux.hpp:57-61)
      if (own_pt)
#pragma unroll
        for (int c = 0; c < 5; ++c) cc[(size_t)c * f] = sE[c][tid] * liftO;
    } else {
      // project_from_mortar: first face dimension (thread = (a, b'), a < N, b' < NM)
      if (tid < N * NM) {
        const int qa = tid % N, b2 = tid / N;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          double v = 0.0;
          for (int m = 0; m < NM; ++m) v += sR[qa * NM + m] * sE[c][m + NM * b2];
          sF[c][tid] = v;
        }
      }
      __syncthreads();
      if (own_pt) {
        const int qa = tid % N, qb = tid / N;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          double v = 0.0;
          for (int m = 0; m < NM; ++m) v += sR[qb * NM + m] * sF[c][qa + N * m];
          cc[(size_t)c * f] = v * liftO;
        }
      }
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------
// ConstraintPreservingBjorhus (both types) on external faces:
// a TimeDerivative-type boundary condition (BoundaryConditionsImpl.hpp:566-670):
// the reference slices the volume time derivative and the volume partial
// derivatives to the face and adds the returned corrections to dt on the face
// points, without lifting.  One CTA per Bjorhus face, one thread per face point:
// the thread forms the 150 logical derivatives at its point from global memory,
// re-evaluates the volume time derivative there with the same per-point code as
// the volume kernel (the volume kernel runs later and adds the corrections in
// the same pass), evaluates the boundary condition (bjorhus.cuh) and writes the
// corrections into the face's corr slots.  kGauge: 0 Harmonic, 1 fields.
// --------------------------------------------------------------------------
struct BjorhusArgs {
  const double* u;
  const double* invjac;
  const double* stat;
  const double* gH;
  const double* gdH;
  const double* coords;
  const double* D;
  double* corr;
  const int32_t* faces;  // [n][3] = element, direction, physical (0/1)
  DampedHarmonicParams dh;
  int elem_begin, elem_end;  // elem_end > 0: only the faces of these elements (lts.cu)
};

// Everything that does not depend on N: volume time derivative at the point,
// the boundary condition's inputs, the condition itself.  One (not inlined) copy.
// d = face direction, dlog[c][jhat] = logical derivatives, corr[50] in Variables
// component order.
static __device__ __noinline__ void gh_bjorhus_point(int gauge, const DampedHarmonicParams& dh,
                                                     bool physical, int d,
                                                     const double (&g)[10],
                                                     const double (&pi)[10],
                                                     const double (&phi)[3][10],
                                                     const double (&J)[3][3], double gamma0,
                                                     double gamma1, double gamma2,
                                                     GaugeH& gh, const double (&x)[3],
                                                     const double (&dlog)[50][3],
                                                     double (&corr)[50]) {
  GhContext ctx;
  double Q[10], dtv[50];
  {
    GaugeInput gin;
    gin.fields = &gh;
    if (gauge == 0) {
      gh_prologue<0>(g, pi, phi, J, gamma0, gamma1, gamma2, gin, ctx, Q);
    } else if (gauge == 1) {
      gh_prologue<1>(g, pi, phi, J, gamma0, gamma1, gamma2, gin, ctx, Q);
    } else {
      // DampedHarmonic: the prologue evaluates H_a and d_a H_b at the point; keep them
      gin.dh = dh;
      for (int xx = 0; xx < 3; ++xx) gin.x[xx] = x[xx];
      gin.computed = &gh;
      gh_prologue<2>(g, pi, phi, J, gamma0, gamma1, gamma2, gin, ctx, Q);
    }
#pragma unroll 1
    for (int s = 0; s < 10; ++s) {
      double ph[3], dph[3][3], oph[3];
      for (int m = 0; m < 3; ++m) {
        ph[m] = phi[m][s];
        for (int xx = 0; xx < 3; ++xx) dph[m][xx] = dlog[20 + m + 3 * s][xx];
      }
      gh_pair_rhs(ctx, Q[s], g[s], pi[s], ph, dlog[s], dlog[10 + s], dph, dtv[s], dtv[10 + s],
                  oph);
      for (int m = 0; m < 3; ++m) dtv[20 + m + 3 * s] = oph[m];
    }
  }
  BjorhusInput in;
  in.physical = physical;
  {
    const double sign = (d & 1) ? 1.0 : -1.0;
    const int dim = d >> 1;
    double unn[3];
    for (int xx = 0; xx < 3; ++xx) unn[xx] = sign * J[dim][xx];
    GhFaceSide sd;
    gh_face_side(g, unn, gamma1, gamma2, sd);
    Geom3p1 q;
    geom_from_metric(g, q);
    in.lapse = q.lapse;
    in.gamma1 = gamma1;
    in.gamma2 = gamma2;
    const double il2 = 1.0 / (q.lapse * q.lapse);
    in.ipsi[0][0] = -il2;
    in.t_up[0] = 1.0 / q.lapse;
    for (int xx = 0; xx < 3; ++xx) {
      in.n_lo[xx] = sd.n_lo[xx];
      in.shift[xx] = q.shift[xx];
      in.x[xx] = x[xx];
      in.ipsi[0][xx + 1] = in.ipsi[xx + 1][0] = q.shift[xx] * il2;
      in.t_up[xx + 1] = -q.shift[xx] / q.lapse;
      for (int y = 0; y < 3; ++y)
        in.ipsi[xx + 1][y + 1] = q.ig[sym3(xx, y)] - q.shift[xx] * q.shift[y] * il2;
    }
#pragma unroll 1
    for (int aa = 0; aa < 4; ++aa) {
      in.H[aa] = gh.H[aa];
#pragma unroll 1
      for (int bb = 0; bb < 4; ++bb) {
        const int s = sym4(aa, bb);
        in.dH[aa][bb] = gh.dH[aa][bb];
        in.g[aa][bb] = g[s];
        in.pi[aa][bb] = pi[s];
        in.dt_g[aa][bb] = dtv[s];
        in.dt_pi[aa][bb] = dtv[10 + s];
        for (int m = 0; m < 3; ++m) {
          in.phi[m][aa][bb] = phi[m][s];
          in.dt_phi[m][aa][bb] = dtv[20 + m + 3 * s];
        }
        // inertial derivatives d_x = J(jhat, x) d_jhat (PartialDerivatives.tpp:79-109)
        for (int xx = 0; xx < 3; ++xx) {
          double dgx = 0.0, dpx = 0.0;
          for (int jh = 0; jh < 3; ++jh) {
            dgx += J[jh][xx] * dlog[s][jh];
            dpx += J[jh][xx] * dlog[10 + s][jh];
          }
          in.c3[xx][aa][bb] = dgx - phi[xx][s];   // three-index constraint
          in.d_pi[xx][aa][bb] = dpx;
          for (int m = 0; m < 3; ++m) {
            double v = 0.0;
            for (int jh = 0; jh < 3; ++jh) v += J[jh][xx] * dlog[20 + m + 3 * s][jh];
            in.d_phi[xx][m][aa][bb] = v;
          }
        }
      }
    }
  }
  BjorhusOutput out;
  bjorhus_constraint_preserving(in, out);
  for (int aa = 0; aa < 4; ++aa)
    for (int bb = aa; bb < 4; ++bb) {
      const int s = sym4(aa, bb);
      corr[s] = out.g[aa][bb];
      corr[10 + s] = out.pi[aa][bb];
      for (int m = 0; m < 3; ++m) corr[20 + m + 3 * s] = out.phi[m][aa][bb];
    }
}

// gauge: 0 Harmonic, 1 gauge fields from memory, 2 DampedHarmonic
template <int N>
__global__ void __launch_bounds__((N * N + 31) / 32 * 32)
    gh_bjorhus_kernel(BjorhusArgs a, int gauge) {
  constexpr int npad = Cfg<N>::npad, f = N * N, T = (N * N + 31) / 32 * 32;
  __shared__ double sD[N * N];
  for (int idx = threadIdx.x; idx < N * N; idx += T) sD[idx] = a.D[idx];
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid >= f) return;
  const int e = a.faces[3 * blockIdx.x], d = a.faces[3 * blockIdx.x + 1];
  if (a.elem_end > 0 && (e < a.elem_begin || e >= a.elem_end)) return;
  const bool physical = a.faces[3 * blockIdx.x + 2] != 0;
  const int qa = tid % N, qb = tid / N;
  const int p = face_point<N>(d, qa, qb);
  const int i = p % N, j = (p / N) % N, k = p / (N * N);
  const double* __restrict__ ue = a.u + (size_t)e * 50 * npad;

  double g[10], pi[10], phi[3][10];
  for (int s = 0; s < 10; ++s) {
    g[s] = __ldg(ue + (size_t)s * npad + p);
    pi[s] = __ldg(ue + (size_t)(10 + s) * npad + p);
    for (int m = 0; m < 3; ++m) phi[m][s] = __ldg(ue + (size_t)(20 + m + 3 * s) * npad + p);
  }
  double J[3][3], x[3];
  const double* je = a.invjac + (size_t)e * 9 * npad + p;
  const double* xe = a.coords + (size_t)e * 3 * npad + p;
  for (int jh = 0; jh < 3; ++jh) {
    x[jh] = __ldg(xe + (size_t)jh * npad);
    for (int xx = 0; xx < 3; ++xx) J[jh][xx] = __ldg(je + (size_t)(jh + 3 * xx) * npad);
  }
  const double* se = a.stat + (size_t)e * 3 * npad + p;
  const double gamma0 = __ldg(se), gamma1 = __ldg(se + npad), gamma2 = __ldg(se + 2 * npad);
  GaugeH gh;
  for (int xx = 0; xx < 4; ++xx) {
    gh.H[xx] = 0.0;
    for (int y = 0; y < 4; ++y) gh.dH[xx][y] = 0.0;
  }
  if (gauge == 1) {
    const double* he = a.gH + (size_t)e * 4 * npad + p;
    const double* dhe = a.gdH + (size_t)e * 16 * npad + p;
    for (int xx = 0; xx < 4; ++xx) {
      gh.H[xx] = __ldg(he + (size_t)xx * npad);
      for (int y = 0; y < 4; ++y) gh.dH[xx][y] = __ldg(dhe + (size_t)(xx + 4 * y) * npad);
    }
  }
  // logical derivatives of all 50 components at the point (K1)
  // (five components per trip: their 15 N loads are independent, so one trip costs
  // about one L2 latency instead of five)
  double dlog[50][3];
#pragma unroll 1
  for (int c0 = 0; c0 < 50; c0 += 5) {
    double acc[5][3];
#pragma unroll
    for (int cc = 0; cc < 5; ++cc) acc[cc][0] = acc[cc][1] = acc[cc][2] = 0.0;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const double di = sD[i * N + m], dj = sD[j * N + m], dk = sD[k * N + m];
#pragma unroll
      for (int cc = 0; cc < 5; ++cc) {
        const double* tc = ue + (size_t)(c0 + cc) * npad;
        acc[cc][0] = fma(di, __ldg(tc + m + N * (j + N * k)), acc[cc][0]);
        acc[cc][1] = fma(dj, __ldg(tc + i + N * (m + N * k)), acc[cc][1]);
        acc[cc][2] = fma(dk, __ldg(tc + i + N * (j + N * m)), acc[cc][2]);
      }
    }
#pragma unroll
    for (int cc = 0; cc < 5; ++cc) {
      dlog[c0 + cc][0] = acc[cc][0];
      dlog[c0 + cc][1] = acc[cc][1];
      dlog[c0 + cc][2] = acc[cc][2];
    }
  }
  double corr[50];
  gh_bjorhus_point(gauge, a.dh, physical, d, g, pi, phi, J, gamma0, gamma1, gamma2, gh, x, dlog,
                   corr);
  double* cf = a.corr + (size_t)e * 10 * 30 * f + (size_t)d * 5 * f + tid;
#pragma unroll 1
  for (int s = 0; s < 10; ++s) {
    double* cs = cf + (size_t)s * 30 * f;
    cs[0] = corr[s];
    cs[(size_t)f] = corr[10 + s];
    for (int m = 0; m < 3; ++m) cs[(size_t)(2 + m) * f] = corr[20 + m + 3 * s];
  }
}

// --------------------------------------------------------------------------
// Halo pack: face slice of u + the static neighbour-side data
// --------------------------------------------------------------------------
struct PackArgs {
  const double* u;
  const double* invjac;
  const double* stat;
  const int32_t* map;  // [G][2] = element, direction
  double* send;        // [G][HC][f]
  int nghost;
};

template <int N, int C>
__global__ void __launch_bounds__(128) pack_halo_kernel(PackArgs a) {
  constexpr int npad = Cfg<N>::npad, f = N * N;
  constexpr int S = (C == 50) ? 3 : 1;       // static comps per element
  constexpr int HC = C + 3 + (C == 50 ? 2 : 1);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.nghost * f) return;
  const int q = (int)(idx % f);
  const int gi = (int)(idx / f);
  const int e = a.map[2 * gi], d = a.map[2 * gi + 1];
  const int p = face_point<N>(d, q % N, q / N);
  const int dim = d >> 1;
  double* out = a.send + (size_t)gi * HC * f + q;
  const double* ue = a.u + (size_t)e * C * npad + p;
#pragma unroll 5
  for (int c = 0; c < C; ++c) out[(size_t)c * f] = __ldg(ue + (size_t)c * npad);
  const double* je = a.invjac + (size_t)e * 9 * npad + p;
#pragma unroll
  for (int x = 0; x < 3; ++x) out[(size_t)(C + x) * f] = __ldg(je + (size_t)(dim + 3 * x) * npad);
  const double* se = a.stat + (size_t)e * S * npad + p;
  if (C == 50) {
    out[(size_t)(C + 3) * f] = __ldg(se + npad);
    out[(size_t)(C + 4) * f] = __ldg(se + 2 * npad);
  } else {
    out[(size_t)(C + 3) * f] = __ldg(se);
  }
}

// --------------------------------------------------------------------------
// u <- a*u + sum_j c_j v_j   (K12+K13: UpdateU / History, flat over the block)
// --------------------------------------------------------------------------
struct LincombArgs {
  double* u;
  double a;
  int nterms;
  double c[8];
  const double* v[8];
  long long len2;  // number of double2
};

static __global__ void __launch_bounds__(256) lincomb_kernel(LincombArgs p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  double2* __restrict__ u2 = reinterpret_cast<double2*>(p.u);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.len2;
       idx += stride) {
    double2 r = u2[idx];
    r.x *= p.a;
    r.y *= p.a;
    for (int j = 0; j < p.nterms; ++j) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(p.v[j]) + idx);
      r.x = fma(p.c[j], v.x, r.x);
      r.y = fma(p.c[j], v.y, r.y);
    }
    u2[idx] = r;
  }
}

// --------------------------------------------------------------------------
// Stand-alone partial_derivatives (operator API, gauge fields): one CTA per
// (element, component): du[(ob + c*oc + i)] = J(jhat, i) d_jhat u_c
// --------------------------------------------------------------------------
struct DerivArgs {
  const double* u;       // [E][C][npad]
  const double* invjac;  // [E][9][npad]
  double* du;            // [E][CO][npad]
  const double* D;
  int C, CO, out_base, out_cstride;
};

template <int N>
__global__ void __launch_bounds__(256) partial_derivatives_kernel(DerivArgs a) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad;
  __shared__ __align__(16) double tile[npad];
  __shared__ double sD[N * N];
  const int e = blockIdx.x / a.C, c = blockIdx.x % a.C;
  const double* uc = a.u + ((size_t)e * a.C + c) * npad;
  for (int p = threadIdx.x; p < n; p += blockDim.x) tile[p] = uc[p];
  for (int p = threadIdx.x; p < N * N; p += blockDim.x) sD[p] = a.D[p];
  __syncthreads();
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const int i = p % N, j = (p / N) % N, k = p / (N * N);
    double Di[N], Dj[N], Dk[N], d[3];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      Di[m] = sD[i * N + m];
      Dj[m] = sD[j * N + m];
      Dk[m] = sD[k * N + m];
    }
    logical_derivs<N>(tile, i, j, k, Di, Dj, Dk, d);
    const double* je = a.invjac + (size_t)e * 9 * npad + p;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      double v = __ldg(je + (size_t)(0 + 3 * x) * npad) * d[0];
      v += __ldg(je + (size_t)(1 + 3 * x) * npad) * d[1];
      v += __ldg(je + (size_t)(2 + 3 * x) * npad) * d[2];
      a.du[((size_t)e * a.CO + a.out_base + c * a.out_cstride + x) * npad + p] = v;
    }
  }
}

// --------------------------------------------------------------------------
// Moving mesh, volume terms (systems without fluxes: VolumeTermsImpl.tpp:155-235,
// dt u += v_g^i d_i u; GH TimeDerivative.cpp:237-300,372-378: gamma1 v_g.C3 in dt g and
// gamma1 gamma2 v_g.C3 in dt Pi, C3_iab = d_i g_ab - Phi_iab).  Runs after the volume kernel
// of a static mesh on its output (a moving mesh is not on the measured path: the fused update
// is off and the derivatives are formed a second time here).  One CTA per (element,
// component); the CTA of Pi_s also differentiates g_s.
// --------------------------------------------------------------------------
struct MeshVelocityArgs {
  const double* u;       // [E][C][npad]
  const double* invjac;  // [E][9][npad]
  const double* stat;    // [E][S][npad] (GH: gamma0, gamma1, gamma2)
  const double* mesh_v;  // [E][3][npad]
  const double* D;
  double* dt;            // [E][C][npad]
  int C, elem_begin, gh;
};

template <int N>
__global__ void __launch_bounds__(256) mesh_velocity_terms_kernel(MeshVelocityArgs a) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad;
  __shared__ __align__(16) double tile[2][npad];
  __shared__ double sD[N * N];
  const int e = a.elem_begin + blockIdx.x / a.C, c = blockIdx.x % a.C;
  const int s = (a.gh && c < 20) ? c % 10 : -1;  // g_s whose C3 enters this component
  const double* ue = a.u + (size_t)e * a.C * npad;
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    tile[0][p] = ue[(size_t)c * npad + p];
    if (s >= 0) tile[1][p] = ue[(size_t)s * npad + p];
  }
  for (int p = threadIdx.x; p < N * N; p += blockDim.x) sD[p] = a.D[p];
  __syncthreads();
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const int i = p % N, j = (p / N) % N, k = p / (N * N);
    double Di[N], Dj[N], Dk[N], d[3], dg[3];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      Di[m] = sD[i * N + m];
      Dj[m] = sD[j * N + m];
      Dk[m] = sD[k * N + m];
    }
    logical_derivs<N>(tile[0], i, j, k, Di, Dj, Dk, d);
    if (s >= 0) logical_derivs<N>(tile[1], i, j, k, Di, Dj, Dk, dg);
    const double* je = a.invjac + (size_t)e * 9 * npad + p;
    const double* ve = a.mesh_v + (size_t)e * 3 * npad + p;
    double adv = 0.0, vc3 = 0.0;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      const double j0 = __ldg(je + (size_t)(0 + 3 * x) * npad);
      const double j1 = __ldg(je + (size_t)(1 + 3 * x) * npad);
      const double j2 = __ldg(je + (size_t)(2 + 3 * x) * npad);
      const double v = __ldg(ve + (size_t)x * npad);
      double du = j0 * d[0];
      du += j1 * d[1];
      du += j2 * d[2];
      adv += v * du;
      if (s >= 0) {
        double dgx = j0 * dg[0];
        dgx += j1 * dg[1];
        dgx += j2 * dg[2];
        vc3 += v * (dgx - __ldg(ue + (size_t)(20 + x + 3 * s) * npad + p));
      }
    }
    double* out = a.dt + ((size_t)e * a.C + c) * npad + p;
    double r = *out;
    if (s >= 0) {
      const double* se = a.stat + (size_t)e * 3 * npad + p;
      const double g1 = __ldg(se + npad);
      r += (c < 10 ? g1 : g1 * __ldg(se + 2 * npad)) * vc3;
    }
    *out = r + adv;
  }
}

// --------------------------------------------------------------------------
// GH constraint diagnostics for the parity norms: sums over all
// points of |C_a|^2 (gauge constraint, Constraints.cpp:965-1000, harmonic or
// field gauge), |C_iab|^2 (three-index, :935-962) and |C_iab|^2 of the
// four-index constraint eps_ijk d_j Phi_kab (:1070-1100), independent components
// only, like ObserveNorms' L2Norm with Components: Sum.  Not on the hot path.
// --------------------------------------------------------------------------
struct ConstraintArgs {
  const double* u;
  const double* invjac;
  const double* gH;  // or nullptr (H = 0)
  const double* D;
  double* sums;      // [3]
};

template <int N>
__global__ void __launch_bounds__(256) gh_constraints_kernel(ConstraintArgs a) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tile = reinterpret_cast<double*>(smem_raw);  // [4][npad]
  double* sD = tile + 4 * npad;                        // [N*N]
  __shared__ double red[3][8];
  const int e = blockIdx.x;
  const double* ue = a.u + (size_t)e * 50 * npad;
  for (int p = threadIdx.x; p < N * N; p += blockDim.x) sD[p] = a.D[p];
  double acc[3] = {0.0, 0.0, 0.0};
  // gauge constraint
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    double g[10], pi[10], phi[3][10], Gam[4];
#pragma unroll
    for (int s = 0; s < 10; ++s) {
      g[s] = ue[(size_t)s * npad + p];
      pi[s] = ue[(size_t)(10 + s) * npad + p];
#pragma unroll
      for (int m = 0; m < 3; ++m) phi[m][s] = ue[(size_t)(20 + m + 3 * s) * npad + p];
    }
    gh_trace_christoffel(g, pi, phi, Gam);
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const double c = Gam[x] + (a.gH ? a.gH[((size_t)e * 4 + x) * npad + p] : 0.0);
      acc[0] += c * c;
    }
  }
  // three- and four-index constraints, pair by pair
  for (int s = 0; s < 10; ++s) {
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
      tile[p] = ue[(size_t)s * npad + p];
#pragma unroll
      for (int m = 0; m < 3; ++m) tile[(1 + m) * npad + p] = ue[(size_t)(20 + m + 3 * s) * npad + p];
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
      const int i = p % N, j = (p / N) % N, k = p / (N * N);
      double Di[N], Dj[N], Dk[N], dl[4][3], J[3][3];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        Di[m] = sD[i * N + m];
        Dj[m] = sD[j * N + m];
        Dk[m] = sD[k * N + m];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) logical_derivs<N>(tile + c * npad, i, j, k, Di, Dj, Dk, dl[c]);
      const double* je = a.invjac + (size_t)e * 9 * npad + p;
#pragma unroll
      for (int jh = 0; jh < 3; ++jh)
#pragma unroll
        for (int x = 0; x < 3; ++x) J[jh][x] = je[(size_t)(jh + 3 * x) * npad];
      double d[4][3];  // inertial derivatives d_x of g_s, Phi_0s, Phi_1s, Phi_2s
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          double v = J[0][x] * dl[c][0];
          v += J[1][x] * dl[c][1];
          v += J[2][x] * dl[c][2];
          d[c][x] = v;
        }
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        const double c3 = d[0][x] - tile[(1 + x) * npad + p];
        acc[1] += c3 * c3;
      }
      const double c40 = d[3][1] - d[2][2];  // eps_0jk d_j Phi_k = d_1 Phi_2 - d_2 Phi_1
      const double c41 = d[1][2] - d[3][0];
      const double c42 = d[2][0] - d[1][1];
      acc[2] += c40 * c40 + c41 * c41 + c42 * c42;
    }
  }
  // block reduction
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    atomicAdd(a.sums + threadIdx.x, v);
  }
}

// --------------------------------------------------------------------------
// AnalyticChristoffel gauge with the GaugeWave solution, evaluated at time t:
// H_a = -Gamma_a[analytic(x, t)]  (AnalyticChristoffel.cpp:76-133,
// GaugeWave.hpp:34-50).  The spatial derivative is then taken numerically by
// partial_derivatives_kernel (:136-143), d_t H_a = 0 (:145-147).
// --------------------------------------------------------------------------
// --------------------------------------------------------------------------
// Exponential filter after the substep: dg::Actions::Filter<
// Filters::Exponential<0>> = apply_matrices(u, {F, F, F}) on every evolved
// component (LinearOperators/ExponentialFilter.cpp:45-76, Spectral/Filtering.cpp
// :20-32).
// --------------------------------------------------------------------------
struct FilterArgs {
  double* u;        // [E][C][npad]
  int ntiles;       // E * C component blocks
  double Fm[144];   // [N][N] row-major filter matrix (kernel parameter = constant bank:
                    // with the loops unrolled every entry is an immediate operand of its DFMA)
};

// Line form: a thread filters one whole grid line in registers -- N loads, N*N FMAs whose
// matrix operand comes from the constant bank, N stores -- so a pass costs one shared-memory
// load and one store per point (round 1: one thread per point, both operands of every FMA
// from shared memory, 8.4 ms for 6144 elements at N = 12; round 2a: R x N register blocks of
// the matrix, 4 loads per point and pass, 2.85 ms, shared-memory wavefronts at 71 %).
// A CTA works on G component blocks at once, N*N threads each.  The thread (a, b) of a block
//   loads the zeta line (i, j) = (a, b) from global memory (coalesced) into the padded tile,
//   filters the xi line (j, k) = (a, b), then the eta line (i, k) = (a, b) in place,
//   filters the zeta line (i, j) = (a, b) and stores it to global memory (coalesced);
// xi, eta, zeta in that order like ApplyMatrices.cpp.  Tile index i + RS j + PS k with an odd
// row stride RS and a plane stride PS = N mod 16: no bank conflicts along xi and eta.
template <int N>
struct FilterCfg {
  static constexpr int f = N * N;
  static constexpr int G = 320 / f > 0 ? 320 / f : 1;  // component blocks per CTA pass
  static constexpr int T = (G * f + 31) / 32 * 32;
  static constexpr int RS = N | 1;
  static constexpr int PS = RS * N + ((N - RS * N) % 16 + 16) % 16;
  static constexpr int tile = PS * N;
  static constexpr int smem_bytes = G * tile * 8;
};

// dst[stride * q] = sum_m F[q][m] x[m]
template <int N>
__device__ __forceinline__ void filter_line(const FilterArgs& a, const double (&x)[N],
                                            double* __restrict__ dst, int stride) {
#pragma unroll
  for (int q = 0; q < N; ++q) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < N; ++m) s = fma(a.Fm[q * N + m], x[m], s);
    dst[stride * q] = s;
  }
}

template <int N>
__global__ void __launch_bounds__(FilterCfg<N>::T) exponential_filter_kernel(
    const __grid_constant__ FilterArgs a) {
  using F = FilterCfg<N>;
  constexpr int npad = Cfg<N>::npad, f = N * N, RS = F::RS, PS = F::PS, G = F::G;
  extern __shared__ __align__(16) unsigned char filter_smem[];
  const int tid = threadIdx.x;
  const int g = tid / f, l = tid % f, la = l % N, lb = l / N;
  double* t = reinterpret_cast<double*>(filter_smem) + g * F::tile;
  const bool mine = tid < G * f && blockIdx.x * G + g < a.ntiles;
  double* ug = a.u + (size_t)(blockIdx.x * G + (mine ? g : 0)) * npad + l;
  double x[N];
  if (mine) {
#pragma unroll
    for (int q = 0; q < N; ++q) x[q] = ug[f * q];
#pragma unroll
    for (int q = 0; q < N; ++q) t[la + RS * lb + PS * q] = x[q];
  }
  __syncthreads();
  if (mine) {  // xi: the line (j, k) = (la, lb)
    double* line = t + RS * la + PS * lb;
#pragma unroll
    for (int m = 0; m < N; ++m) x[m] = line[m];
    filter_line<N>(a, x, line, 1);
  }
  __syncthreads();
  if (mine) {  // eta: the line (i, k) = (la, lb)
    double* line = t + la + PS * lb;
#pragma unroll
    for (int m = 0; m < N; ++m) x[m] = line[RS * m];
    filter_line<N>(a, x, line, RS);
  }
  __syncthreads();
  if (mine) {  // zeta: the line (i, j) = (la, lb), straight to global memory
    const double* line = t + la + RS * lb;
#pragma unroll
    for (int m = 0; m < N; ++m) x[m] = line[PS * m];
    filter_line<N>(a, x, ug, f);
  }
}

// AnalyticChristoffel for a static analytic solution (AnalyticChristoffel.cpp:
// 76-147): H_a = -Gamma_a of the analytic (g, Pi, Phi); the spatial derivative
// is then taken by partial_derivatives_kernel, d_t H_a = 0.
struct GaugeFromStateArgs {
  const double* u;  // analytic state [E][50][npad]
  double* gH;       // [E][4][npad]
  int nelem;
};

template <int N>
__global__ void __launch_bounds__(256) gauge_h_from_state_kernel(GaugeFromStateArgs a) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.nelem * n) return;
  const int e = (int)(idx / n), p = (int)(idx % n);
  const double* ue = a.u + (size_t)e * 50 * npad + p;
  double g[10], pi[10], phi[3][10], Gam[4];
#pragma unroll
  for (int s = 0; s < 10; ++s) {
    g[s] = ue[(size_t)s * npad];
    pi[s] = ue[(size_t)(10 + s) * npad];
#pragma unroll
    for (int m = 0; m < 3; ++m) phi[m][s] = ue[(size_t)(20 + m + 3 * s) * npad];
  }
  gh_trace_christoffel(g, pi, phi, Gam);
#pragma unroll
  for (int x = 0; x < 4; ++x) a.gH[((size_t)e * 4 + x) * npad + p] = -Gam[x];
}

struct GaugeWaveArgs {
  const double* coords;  // [E][3][npad]
  double* gH;            // [E][4][npad]
  double amplitude, wavelength, time;
  int nelem;
};

template <int N>
__global__ void __launch_bounds__(256) gauge_wave_h_kernel(GaugeWaveArgs a) {
  constexpr int n = Cfg<N>::n, npad = Cfg<N>::npad;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.nelem * n) return;
  const int e = (int)(idx / n), p = (int)(idx % n);
  const double x = a.coords[((size_t)e * 3) * npad + p];
  const double omega = 2.0 * M_PI / a.wavelength;
  const double H = 1.0 - a.amplitude * sin(omega * (x - a.time));
  const double dH = -omega * a.amplitude * cos(omega * (x - a.time));
  // g = diag(-H, H, 1, 1); d_t g_00 = dH, d_t g_11 = -dH, d_x g_00 = -dH,
  // d_x g_11 = dH.  Gamma_a = g^{bc} Gamma_a,bc with g^{00} = -1/H, g^{11}=1/H
  //   Gamma_0 = g^{00} Gamma_0,00 + g^{11} Gamma_0,11
  //           = (-1/H)(1/2 d_t g_00) + (1/H)(d_x g_01.. - 1/2 d_t g_11)
  const double G00 = -1.0 / H, G11 = 1.0 / H;
  const double chr_0_00 = 0.5 * dH;             // 1/2 d_t g_00
  const double chr_0_11 = -0.5 * (-dH);         // -1/2 d_t g_11
  const double chr_1_00 = -0.5 * (-dH);         // -1/2 d_x g_00
  const double chr_1_11 = 0.5 * dH;             // 1/2 d_x g_11
  const double gam0 = G00 * chr_0_00 + G11 * chr_0_11;
  const double gam1 = G00 * chr_1_00 + G11 * chr_1_11;
  double* h = a.gH + (size_t)e * 4 * npad + p;
  h[0] = -gam0;
  h[(size_t)npad] = -gam1;
  h[(size_t)2 * npad] = 0.0;
  h[(size_t)3 * npad] = 0.0;
}

}  // namespace dg
