// microbench: bandwidth of reading a [R][P] float matrix in column chunks of CHUNK bytes through all R rows
// (the access pattern of a fold work item: every energy plane of a few pixels), LDG.128 vs chunk width.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template <int CHUNK, int MLP>
__global__ void __launch_bounds__(256) stride_read(const float* __restrict__ a, int R, long long P, int n_chunks, float* out, unsigned* counter) {
  constexpr int LPR = CHUNK / 16;       // lanes per row
  constexpr int RPI = 32 / LPR;         // rows per warp instruction
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (int c = warp_global; c < n_chunks; c += n_warps) {
    const long long col = (long long)c * (CHUNK / 4) + (lane % LPR) * 4;
    const int r0 = lane / LPR;
    for (int rb = 0; rb < R; rb += RPI * MLP) {
      float4 v[MLP];
#pragma unroll
      for (int m = 0; m < MLP; ++m) {
        const int r = rb + m * RPI + r0;
        v[m] = make_float4(0, 0, 0, 0);
        if (r < R) v[m] = __ldg(reinterpret_cast<const float4*>(a + (long long)r * P + col));
      }
#pragma unroll
      for (int m = 0; m < MLP; ++m) acc += v[m].x + v[m].y + v[m].z + v[m].w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
template <int CHUNK, int MLP> void run(const float* a, int R, long long P, float* out, int blocks_per_sm) {
  const int n_chunks = (int)(P * 4 / CHUNK);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * blocks_per_sm;
  stride_read<CHUNK, MLP><<<grid, 256>>>(a, R, P, n_chunks, out, nullptr);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) stride_read<CHUNK, MLP><<<grid, 256>>>(a, R, P, n_chunks, out, nullptr);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("chunk %4d B  mlp %2d  blocks/SM %d : %.3f ms  %.0f GB/s  (%s)\n", CHUNK, MLP, blocks_per_sm, ms, (double)R * P * 4 / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int R = 60; const long long P = 1000000;
  float* a; cudaMalloc(&a, (size_t)R * P * 4 + 4096); cudaMemset(a, 0, (size_t)R * P * 4 + 4096);
  float* out; cudaMalloc(&out, 16);
  for (int b = 1; b <= 4; b *= 2) {
    run<128, 4>(a, R, P, out, b); run<128, 8>(a, R, P, out, b); run<128, 15>(a, R, P, out, b);
    run<256, 8>(a, R, P, out, b);
    run<512, 4>(a, R, P, out, b); run<512, 8>(a, R, P, out, b); run<512, 15>(a, R, P, out, b);
  }
  // sequential baseline: same kernel with R = 1 row of R*P floats
  run<512, 8>(a, 1, (long long)R * P, out, 4);
  return 0;
}
