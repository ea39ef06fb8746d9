// development probe: throughput of FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a2.x, b2.x); acc[i].y = fmaf(acc[i].y, a2.y, b2.y); }
      else acc[i] = __ffma2_rn(acc[i], a2, b2);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 14;
  for (int wps = 1; wps <= 8; wps *= 2)
  for (int mode = 0; mode < 2; ++mode) {
    dim3 grid(148), block(128 * wps);   // wps warps per scheduler
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, block>>>(out, 0.999f, 0.001f, iters); else k<1><<<grid, block>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * block.x * 16.0 * iters;
    printf("%s warps/sched %d: %.3f ms  %.1f GFMA/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", mode ? "FFMA2" : "FFMA ", wps, ms, fma / ms / 1e6, fma / ms / 1e6 / 148 / 1.9);
  }
  return 0;
}
