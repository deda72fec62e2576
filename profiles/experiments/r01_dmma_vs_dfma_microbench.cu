// microbenchmark: fp64 mma.sync (DMMA) vs DFMA throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a0, double b0) {
  double c[8][2];
  double f[8];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; f[i] = i + threadIdx.x; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[i], a, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int warps_per_sm) {
  double* out; cudaMalloc(&out, 148 * 1024 * 8 * 8);
  int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warps = (double)blocks * threads / 32;
  double dmma = (MODE != 1) ? warps * iters * 8.0 : 0;      // warp-level DMMAs
  double dfma = (MODE != 0) ? warps * iters * 64.0 : 0;     // warp-level DFMAs
  double flops = (dmma * 256 + dfma * 32) * 2;
  printf("%s warps/SM %d: %.3f ms, %.2f TFLOP/s; cycles/DMMA/SMSP %.2f cycles/DFMA/SMSP %.2f (at 1.9GHz)\n",
         name, warps_per_sm, ms, flops / ms * 1e-9,
         dmma ? ms * 1e-3 * 1.9e9 / (dmma / (148 * 4)) : 0.0,
         dfma ? ms * 1e-3 * 1.9e9 / (dfma / (148 * 4)) : 0.0);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16}) {
    run<0>("DMMA only ", w);
    run<1>("DFMA only ", w);
    run<2>("DMMA+DFMA ", w);
  }
  return 0;
}
