#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
// microbench: DMMA m8n8k4 / m16n8k8(if supported) / DFMA / F2F throughput per SM
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#ifdef BIG
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
   : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
   : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
#endif
template <int MODE>
__global__ void __launch_bounds__(256) bench(double* out, int iters, double seed) {
  double acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = seed * (threadIdx.x + i);
  double a = seed + threadIdx.x, b = seed * 0.5;
  float fa[8];
  for (int i = 0; i < 8; ++i) fa[i] = (float)(seed * i + threadIdx.x);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(acc[2 * i], acc[2 * i + 1], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(a, b, acc[i]);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[i] += (double)fa[i]; fa[i] = (float)acc[i + 8] ; }
    }
#ifdef BIG
    else if (MODE == 3) {
      double av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
      for (int i = 0; i < 4; ++i) dmma1688(acc + 4 * i, av, bv);
    } else if (MODE == 4) {
      double av[8] = {a, a + 1, a + 2, a + 3, a+4,a+5,a+6,a+7}, bv[4] = {b, b + 1, b+2, b+3};
#pragma unroll
      for (int i = 0; i < 4; ++i) dmma16816(acc + 4 * i, av, bv);
    }
#endif
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  for (int i = 0; i < 8; ++i) s += fa[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double fma_per_iter_per_warp, int warps_per_sm_blocks) {
  int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int blocks = p.multiProcessorCount * warps_per_sm_blocks;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * 256);
  int iters = 20000;
  bench<MODE><<<blocks, 256>>>(out, 100, 1.0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<MODE><<<blocks, 256>>>(out, iters, 1.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double total = fma_per_iter_per_warp * iters * 8.0 * blocks;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  printf("%-12s blocks/SM %d: %.3f ms  %.2f Gop/s  = %.1f op/clk/SM at %d MHz (nominal)  err=%s\n", name, warps_per_sm_blocks, ms, total / ms * 1e-6,
         total / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  for (int b = 1; b <= 4; b *= 2) {
    run<0>("dmma884", 8 * 256.0, b);
    run<1>("dfma", 16 * 32.0, b);
    run<2>("f2f+dadd", 8 * 32.0, b);
#ifdef BIG
    run<3>("dmma1688", 4 * 1024.0, b);
    run<4>("dmma16816", 4 * 2048.0, b);
#endif
  }
  return 0;
}
