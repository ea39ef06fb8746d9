// development probe: 3-D tiled TMA load (cp.async.bulk.tensor.3d) of a planar fp32 image box with halo
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int BW, int BH>
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int x0, int y0, int c0, int* status) {
  extern __shared__ __align__(128) float tile[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(tile + 3 * BH * BW);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(3u * BH * BW * 4u) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s32(tile)), "l"(&map), "r"(x0), "r"(y0), "r"(c0), "r"(s32(bar)) : "memory");
  }
  unsigned done = 0, spins = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(s32(bar)) : "memory");
    if (!done && ++spins > (1u << 22)) { if (threadIdx.x == 0) *status = -1; return; }
  }
  for (int i = threadIdx.x; i < 3 * BH * BW; i += blockDim.x) out[i] = tile[i];
  if (threadIdx.x == 0) *status = (int)spins + 1;
}
template <int BW, int BH>
void run(enc_fn enc, float* d, int B, int H, int W, int x0, int y0) {
  CUtensorMap m;
  const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {BW, BH, 3};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  float* out; int* st; cudaMalloc(&out, 3 * BH * BW * 4); cudaMalloc(&st, 4); cudaMemset(st, 0, 4);
  const int smem = 3 * BH * BW * 4 + 16;
  cudaFuncSetAttribute(k<BW, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<BW, BH><<<1, 128, smem>>>(m, out, x0, y0, 3, st);
  cudaError_t e = cudaDeviceSynchronize();
  int hs = 0; cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
  std::vector<float> h(3 * BH * BW); cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < 3; ++c) for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
    const int gx = x0 + x, gy = y0 + y, gc = 3 + c;
    const float want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : (float)((gc * H + gy) * W + gx);
    if (h[(c * BH + y) * BW + x] != want) ++bad;
  }
  printf("box %dx%d at (%d,%d): encode %d, sync %s, status %d, mismatches %d\n", BW, BH, x0, y0, (int)r, cudaGetErrorString(e), hs, bad);
  cudaFree(out); cudaFree(st);
}
int main(int argc, char** argv) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc_fn enc = (enc_fn)p;
  printf("entry point %p query %d\n", p, (int)q);
  const int B = 2, H = 192, W = 640;
  std::vector<float> h((size_t)B * 3 * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  switch (which) {
    case 0: run<128, 16>(enc, d, B, H, W, 0, 0); run<128, 16>(enc, d, B, H, W, -1, -1); break;
    case 1: run<136, 18>(enc, d, B, H, W, 124, 15); break;
    case 9: run<136, 18>(enc, d, B, H, W, -4, -1); break;
    case 10: run<136, 18>(enc, d, B, H, W, 508, 175); break;
    case 2: run<128, 18>(enc, d, B, H, W, 127, 15); break;
    case 3: run<144, 16>(enc, d, B, H, W, 127, 15); break;
    case 4: run<160, 16>(enc, d, B, H, W, 127, 15); break;
    case 5: run<136, 18>(enc, d, B, H, W, 127, 15); break;
    case 6: run<192, 18>(enc, d, B, H, W, 127, 15); break;
    case 7: run<256, 18>(enc, d, B, H, W, 127, 15); break;
    case 8: run<128, 20>(enc, d, B, H, W, -1, -1); break;
  }
  return 0;
}
