// md2_monitor.cu - monitoring path: Trainer.compute_depth_losses (/root/reference/trainer.py:498-526)
// with compute_depth_errors (/root/reference/layers.py:251-269) as a handful of launches, no host synchronisation:
//   pred = clamp(bilinear_up(depth, (Hg, Wg), align_corners=False), 1e-3, 80)         trainer.py:504-507
//   mask = (gt > 0) & crop[y0:y1, x0:x1]                                              trainer.py:509-515
//   pred *= median(gt[mask]) / median(pred[mask]);  pred = clamp(pred, 1e-3, 80)      trainer.py:517-521
//   abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 over the masked pixels of the whole batch
// torch.median returns the LOWER median (element (n-1)/2 of the sorted values); it is found here exactly by a 4-pass
// radix select over the float bit patterns (all values are positive, so unsigned order == float order).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/md2_loss.h"

namespace {

constexpr int kT = 256;

struct MonArgs {
  const float* depth;   // (B,1,H,W)
  const float* gt;      // (B,1,Hg,Wg)
  float* pred;          // scratch (B,Hg,Wg): up-sampled, clamped prediction
  unsigned* hist;       // scratch: [2][256] histograms (gt, pred)
  unsigned* state;      // scratch: [0] N, [1..2] prefix (gt, pred), [3..4] remaining rank (gt, pred)
  double* sums;         // scratch: 7 doubles
  float* out;           // 7 floats
  int B, H, W, Hg, Wg, y0, y1, x0, x1;
};

__device__ __forceinline__ bool masked(const MonArgs& a, int y, int x, float g) {
  return g > 0.f && y >= a.y0 && y < a.y1 && x >= a.x0 && x < a.x1;
}

// torch upsample_bilinear2d, align_corners=False
__device__ __forceinline__ float up_bilinear(const float* d, int H, int W, int Hg, int Wg, int y, int x) {
  const float ry = (float)H / (float)Hg, rx = (float)W / (float)Wg;
  float sy = fmaf(ry, (float)y + 0.5f, -0.5f); sy = sy < 0.f ? 0.f : sy;
  float sx = fmaf(rx, (float)x + 0.5f, -0.5f); sx = sx < 0.f ? 0.f : sx;
  const int yy0 = (int)sy, xx0 = (int)sx;
  const int yy1 = yy0 + (yy0 < H - 1 ? 1 : 0), xx1 = xx0 + (xx0 < W - 1 ? 1 : 0);
  const float ly1 = sy - (float)yy0, ly0 = 1.f - ly1, lx1 = sx - (float)xx0, lx0 = 1.f - lx1;
  return ly0 * (lx0 * __ldg(d + yy0 * W + xx0) + lx1 * __ldg(d + yy0 * W + xx1)) +
         ly1 * (lx0 * __ldg(d + yy1 * W + xx0) + lx1 * __ldg(d + yy1 * W + xx1));
}

__global__ void k_mon_reset(MonArgs a) {
  const int i = threadIdx.x;
  if (i < 512) a.hist[i] = 0u;
  if (i < 8) a.state[i] = 0u;
  if (i < 7) a.sums[i] = 0.0;
}

// pass over the masked pixels: up-sample + clamp, count
__global__ void __launch_bounds__(kT) k_mon_prepare(MonArgs a) {
  const long long n = (long long)a.B * a.Hg * a.Wg;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned c = 0;
  if (i < n) {
    const int plane = a.Hg * a.Wg;
    const int b = (int)(i / plane), p = (int)(i - (long long)b * plane);
    const int y = p / a.Wg, x = p - y * a.Wg;
    if (masked(a, y, x, __ldg(a.gt + i))) {
      float v = up_bilinear(a.depth + (size_t)b * a.H * a.W, a.H, a.W, a.Hg, a.Wg, y, x);
      v = fminf(fmaxf(v, 1e-3f), 80.f);
      a.pred[i] = v;
      c = 1;
    }
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(a.state, c);
}

// radix-select pass `pass` (most significant byte first): histogram of byte (3 - pass) among the values whose higher
// bytes equal the prefix found so far
__global__ void __launch_bounds__(kT) k_mon_hist(MonArgs a, int pass) {
  __shared__ unsigned h[2][256];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) (&h[0][0])[i] = 0u;
  __syncthreads();
  const long long n = (long long)a.B * a.Hg * a.Wg;
  const int plane = a.Hg * a.Wg;
  const int shift = 8 * (3 - pass);
  const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  const unsigned pg = a.state[1], pp = a.state[2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % plane);
    const int y = p / a.Wg, x = p - y * a.Wg;
    const float g = __ldg(a.gt + i);
    if (!masked(a, y, x, g)) continue;
    const unsigned ug = __float_as_uint(g), up = __float_as_uint(a.pred[i]);
    if ((ug & himask) == pg) atomicAdd(&h[0][(ug >> shift) & 255u], 1u);
    if ((up & himask) == pp) atomicAdd(&h[1][(up >> shift) & 255u], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const unsigned v = (&h[0][0])[i];
    if (v) atomicAdd(a.hist + i, v);
  }
}

// picks the bin holding the wanted rank, extends the prefix, clears the histograms
__global__ void k_mon_pick(MonArgs a, int pass) {
  const int which = threadIdx.x;          // 0: gt, 1: pred
  if (which < 2) {
    const unsigned N = a.state[0];
    unsigned rank = pass == 0 ? (N ? (N - 1) / 2 : 0) : a.state[3 + which];
    const unsigned* h = a.hist + which * 256;
    unsigned acc = 0;
    int bin = 0;
    for (; bin < 256; ++bin) {
      if (acc + h[bin] > rank) break;
      acc += h[bin];
    }
    if (bin > 255) bin = 255;
    a.state[3 + which] = rank - acc;
    a.state[1 + which] |= ((unsigned)bin) << (8 * (3 - pass));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) a.hist[i] = 0u;
}

__global__ void __launch_bounds__(kT) k_mon_errors(MonArgs a) {
  const float med_g = __uint_as_float(a.state[1]), med_p = __uint_as_float(a.state[2]);
  const float ratio = med_g / med_p;                                     // trainer.py:519
  const long long n = (long long)a.B * a.Hg * a.Wg;
  const int plane = a.Hg * a.Wg;
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % plane);
    const int y = p / a.Wg, x = p - y * a.Wg;
    const float g = __ldg(a.gt + i);
    if (!masked(a, y, x, g)) continue;
    const float pr = fminf(fmaxf(a.pred[i] * ratio, 1e-3f), 80.f);       // trainer.py:519-521
    const float th = fmaxf(g / pr, pr / g);                              // layers.py:254-268
    const float d = g - pr, dl = logf(g) - logf(pr);
    s[0] += (double)(fabsf(d) / g);
    s[1] += (double)(d * d / g);
    s[2] += (double)(d * d);
    s[3] += (double)(dl * dl);
    s[4] += th < 1.25f ? 1.0 : 0.0;
    s[5] += th < 1.25f * 1.25f ? 1.0 : 0.0;
    s[6] += th < 1.25f * 1.25f * 1.25f ? 1.0 : 0.0;
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    double v = s[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(a.sums + k, v);
  }
}

__global__ void k_mon_final(MonArgs a) {
  const int k = threadIdx.x;
  if (k >= 7) return;
  const double N = (double)a.state[0];
  double m = N > 0 ? a.sums[k] / N : NAN;
  if (k == 2 || k == 3) m = sqrt(m);
  a.out[k] = (float)m;
}

}  // namespace

extern "C" {

int md2_depth_metrics_scratch_bytes(int batch, int gt_height, int gt_width, size_t* bytes) {
  if (!bytes || batch < 1 || gt_height < 1 || gt_width < 1) return MD2_ERR_INVALID_ARGUMENT;
  *bytes = (size_t)batch * gt_height * gt_width * sizeof(float) + 512 * sizeof(unsigned) + 8 * sizeof(unsigned) + 8 * sizeof(double) + 64;
  return MD2_OK;
}

int md2_depth_metrics(const float* depth, const float* depth_gt, float* metrics, void* scratch, size_t scratch_bytes,
                      int batch, int height, int width, int gt_height, int gt_width, int crop_y0, int crop_y1,
                      int crop_x0, int crop_x1, void* stream) {
  size_t need = 0;
  if (md2_depth_metrics_scratch_bytes(batch, gt_height, gt_width, &need) != MD2_OK || !depth || !depth_gt || !metrics ||
      height < 1 || width < 1)
    return MD2_ERR_INVALID_ARGUMENT;
  if (!scratch || scratch_bytes < need) return MD2_ERR_WORKSPACE_TOO_SMALL;
  MonArgs a;
  char* ws = (char*)scratch;
  a.sums = (double*)ws; ws += 8 * sizeof(double);
  a.hist = (unsigned*)ws; ws += 512 * sizeof(unsigned);
  a.state = (unsigned*)ws; ws += 8 * sizeof(unsigned);
  ws = (char*)(((uintptr_t)ws + 15) & ~(uintptr_t)15);
  a.pred = (float*)ws;
  a.depth = depth; a.gt = depth_gt; a.out = metrics;
  a.B = batch; a.H = height; a.W = width; a.Hg = gt_height; a.Wg = gt_width;
  a.y0 = crop_y0 < 0 ? 0 : crop_y0; a.y1 = crop_y1 > gt_height ? gt_height : crop_y1;
  a.x0 = crop_x0 < 0 ? 0 : crop_x0; a.x1 = crop_x1 > gt_width ? gt_width : crop_x1;
  cudaStream_t s = (cudaStream_t)stream;
  const long long n = (long long)batch * gt_height * gt_width;
  const unsigned full = (unsigned)((n + kT - 1) / kT);
  const unsigned grid = full < 148u * 8u ? full : 148u * 8u;
  k_mon_reset<<<1, 512, 0, s>>>(a);
  k_mon_prepare<<<full, kT, 0, s>>>(a);
  for (int pass = 0; pass < 4; ++pass) {
    k_mon_hist<<<grid, kT, 0, s>>>(a, pass);
    k_mon_pick<<<1, 256, 0, s>>>(a, pass);
  }
  k_mon_errors<<<grid, kT, 0, s>>>(a);
  k_mon_final<<<1, 32, 0, s>>>(a);
  return cudaGetLastError() == cudaSuccess ? MD2_OK : MD2_ERR_CUDA;
}

}  // extern "C"
