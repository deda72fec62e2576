// micro-benchmark of the GH volume kernels on synthetic data (experiments only)
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../spectre_b200/csrc/volume_pencil.cuh"
#ifndef NN
#define NN 12
#endif
using namespace dg;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

static std::vector<double> diffmat(int N);

template <typename K>
float run(K k, int blocks, int threads, int smem, GhVolArgs a, const char* name, int reps = 5) {
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads, smem>>>(a);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) k<<<blocks, threads, smem>>>(a);
  cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-44s %8.3f ms\n", name, ms / reps);
  return ms / reps;
}

int main(int argc, char** argv) {
  constexpr int N = NN;
  const int E = argc > 1 ? atoi(argv[1]) : 2048;
  const int npad = Cfg<N>::npad, f = N * N;
  size_t len = (size_t)E * 50 * npad;
  std::vector<double> hu(len);
  // Minkowski + noise: g = diag(-1,1,1,1)
  srand(1);
  for (int e = 0; e < E; ++e)
    for (int c = 0; c < 50; ++c)
      for (int p = 0; p < npad; ++p) {
        double base = 0.0;
        if (c == 0) base = -1.0;
        if (c == 4 || c == 7 || c == 9) base = 1.0;
        hu[((size_t)e * 50 + c) * npad + p] = base + 1e-3 * (rand() / (double)RAND_MAX - 0.5);
      }
  double *u, *dt, *un, *v0, *v1, *J, *st, *corr, *gH, *gdH, *D;
  CK(cudaMalloc(&u, len * 8)); CK(cudaMalloc(&dt, len * 8)); CK(cudaMalloc(&un, len * 8));
  CK(cudaMalloc(&v0, len * 8)); CK(cudaMalloc(&v1, len * 8));
  CK(cudaMemcpy(u, hu.data(), len * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(v0, 0, len * 8)); CK(cudaMemset(v1, 0, len * 8));
  std::vector<double> hJ((size_t)E * 9 * npad, 0.0);
  for (int e = 0; e < E; ++e) for (int c = 0; c < 9; c += 4) for (int p = 0; p < npad; ++p) hJ[((size_t)e * 9 + c) * npad + p] = 2.0;
  CK(cudaMalloc(&J, hJ.size() * 8)); CK(cudaMemcpy(J, hJ.data(), hJ.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&st, (size_t)E * 3 * npad * 8)); CK(cudaMemset(st, 0, (size_t)E * 3 * npad * 8));
  CK(cudaMalloc(&corr, (size_t)E * 300 * f * 8)); CK(cudaMemset(corr, 0, (size_t)E * 300 * f * 8));
  CK(cudaMalloc(&gH, (size_t)E * 4 * npad * 8)); CK(cudaMemset(gH, 0, (size_t)E * 4 * npad * 8));
  CK(cudaMalloc(&gdH, (size_t)E * 16 * npad * 8)); CK(cudaMemset(gdH, 0, (size_t)E * 16 * npad * 8));
  std::vector<double> hD = diffmat(N);
  CK(cudaMalloc(&D, N * N * 8)); CK(cudaMemcpy(D, hD.data(), N * N * 8, cudaMemcpyHostToDevice));
  std::vector<double> all(13 * 144, 0.0);
  for (int i = 0; i < N * N; ++i) all[N * 144 + i] = hD[i];
  CK(cudaMemcpyToSymbol(dgc_diff_matrices, all.data(), all.size() * 8));
  GhVolArgs a{u, dt, J, st, corr, gH, gdH, D, nullptr, {}, 0, {}};
  a.upd.u_new = un; a.upd.a = 1.0; a.upd.c_new = 1e-4; a.upd.nterms = 2;
  a.upd.c[0] = 1e-4; a.upd.c[1] = 1e-4; a.upd.v[0] = v0; a.upd.v[1] = v1;
  GhVolArgs b = a; b.upd.u_new = nullptr;   // no fused update
  GhVolArgs c = a; c.corr = nullptr;        // no corrections
  using P = PCfg<N>;
  printf("N=%d E=%d  pencil: T=%d LPC=%d nchunk=%d stages=%d smem=%d | old: T=%d nchunk=%d stages=%d\n", N, E, P::T, P::LPC,
         P::nchunk, P::nstage, P::smem_bytes, Cfg<N>::T, Cfg<N>::nchunk, Cfg<N>::nstage);
  run(gh_volume_kernel<N, 1>, E * Cfg<N>::nchunk, Cfg<N>::T, gh_volume_smem_bytes<N>(), a, "old fused");
  run(gh_volume_kernel<N, 1>, E * Cfg<N>::nchunk, Cfg<N>::T, gh_volume_smem_bytes<N>(), b, "old no-update");
  run(gh_volume_kernel<N, 1>, E * Cfg<N>::nchunk, Cfg<N>::T, gh_volume_smem_bytes<N>(), c, "old no-corr");
#define PRUN(dbg, args, name) run(gh_volume_pencil_kernel<N, 1, dbg>, E * P::nchunk, P::T, P::smem_bytes, args, name)
  PRUN(0, a, "pencil fused");
  PRUN(0, b, "pencil no-update");
  PRUN(0, c, "pencil no-corr");
  PRUN(8, a, "pencil skip corr loads");
  PRUN(1, a, "pencil skip xi tasks");
  PRUN(2, a, "pencil skip eta tasks");
  PRUN(3, a, "pencil skip xi+eta tasks");
  PRUN(4, a, "pencil skip zeta");
  PRUN(7, a, "pencil skip xi+eta+zeta");
  PRUN(32, a, "pencil skip pair rhs");
  PRUN(64, a, "pencil skip prologue");
  PRUN(64 + 32 + 7, a, "pencil skip all compute");
  PRUN(64 + 32 + 7 + 8, b, "pencil skip all compute+corr+update");
  // correctness: pencil vs old on the same data (dt and u_new), with nonzero corrections
  {
    std::vector<double> hc((size_t)E * 300 * f);
    for (auto& x : hc) x = 1e-2 * (rand() / (double)RAND_MAX - 0.5);
    CK(cudaMemcpy(corr, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice));
    std::vector<double> r0(len), r1(len), w0(len), w1(len);
    CK(cudaMemset(dt, 0, len * 8)); CK(cudaMemset(un, 0, len * 8));
    gh_volume_kernel<N, 1><<<E * Cfg<N>::nchunk, Cfg<N>::T, gh_volume_smem_bytes<N>()>>>(a);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(r0.data(), dt, len * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(w0.data(), un, len * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemset(dt, 0, len * 8)); CK(cudaMemset(un, 0, len * 8));
    gh_volume_pencil_kernel<N, 1, 0><<<E * P::nchunk, P::T, P::smem_bytes>>>(a);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(r1.data(), dt, len * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(w1.data(), un, len * 8, cudaMemcpyDeviceToHost));
    double md = 0, mu = 0, mx = 0; size_t bad = 0, first = (size_t)-1;
    for (int e = 0; e < E; ++e) for (int cc = 0; cc < 50; ++cc) for (int p = 0; p < N * N * N; ++p) {
      size_t i = ((size_t)e * 50 + cc) * npad + p;
      double d = fabs(r0[i] - r1[i]); if (d > md) md = d; if (fabs(r0[i]) > mx) mx = fabs(r0[i]);
      if (d > 1e-9) { ++bad; if (first == (size_t)-1) first = i; }
      d = fabs(w0[i] - w1[i]); if (d > mu) mu = d;
    }
    printf("compare pencil vs old: max|d dt| = %.3e (max|dt| %.3e)  max|d u_new| = %.3e  bad = %zu\n", md, mx, mu, bad);
    if (first != (size_t)-1) {
      size_t e = first / ((size_t)50 * npad), cc = (first / npad) % 50, p = first % npad;
      printf("first bad: element %zu comp %zu point %zu (i %zu j %zu k %zu): old %.6e new %.6e\n", e, cc, p, p % N, (p / N) % N, p / (N * N), r0[first], r1[first]);
      // histogram of bad entries by component and by element
      std::vector<size_t> byc(50, 0); size_t nbe = 0, laste = (size_t)-1;
      for (int e2 = 0; e2 < E; ++e2) { bool any = false; for (int cc2 = 0; cc2 < 50; ++cc2) for (int p2 = 0; p2 < N * N * N; ++p2) { size_t i = ((size_t)e2 * 50 + cc2) * npad + p2; if (fabs(r0[i] - r1[i]) > 1e-9) { byc[cc2]++; any = true; } } if (any) { ++nbe; laste = e2; } }
      printf("bad elements: %zu of %d (last %zu); by comp:", nbe, E, laste);
      for (int cc2 = 0; cc2 < 50; ++cc2) printf(" %zu", byc[cc2]);
      printf("\n");
    }
  }
  return 0;
}

#include <cmath>
static std::vector<double> diffmat(int N) {
  // LGL nodes by Newton on (1-x^2) P'_{N-1}; barycentric differentiation matrix
  int n = N - 1;
  std::vector<double> x(N), w(N), D(N * N);
  for (int i = 0; i < N; ++i) {
    double xi = -cos(M_PI * i / n);
    for (int it = 0; it < 100; ++it) {
      double p0 = 1, p1 = xi;
      for (int k = 2; k <= n; ++k) { double p2 = ((2 * k - 1) * xi * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
      // p1 = P_n, p0 = P_{n-1}; q = (1-x^2) P_n' = n (P_{n-1} - x P_n)
      double q = n * (p0 - xi * p1);
      double dq = -n * (n + 1) * p1;
      if (i == 0 || i == n) break;
      double dx = q / dq; xi -= dx; if (fabs(dx) < 1e-15) break;
    }
    x[i] = (i == 0) ? -1 : (i == n ? 1 : xi);
  }
  for (int i = 0; i < N; ++i) { w[i] = 1; for (int j = 0; j < N; ++j) if (j != i) w[i] /= (x[i] - x[j]); }
  for (int i = 0; i < N; ++i) { double s = 0; for (int j = 0; j < N; ++j) if (j != i) { D[i * N + j] = (w[j] / w[i]) / (x[i] - x[j]); s += D[i * N + j]; } D[i * N + i] = -s; }
  return D;
}
