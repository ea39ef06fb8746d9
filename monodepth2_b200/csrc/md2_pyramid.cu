// md2_pyramid.cu - colour pyramid on the GPU: Pillow's uint8 LANCZOS resize, byte for byte.
//
// The reference builds ("color", f, s) on the host, per frame and per sample, with torchvision
// transforms.Resize(..., interpolation=Image.ANTIALIAS) on PIL images: scale 0 from the native frame,
// scale i from scale i-1 (/root/reference/datasets/mono_dataset.py:57,82-86,98-103), i.e.
// PIL.Image.resize(size, LANCZOS).  Pillow (src/libImaging/Resample.c) does a separable two-pass resample in
// 8-bit fixed point: precompute_coeffs (double), normalize_coeffs_8bpc (22-bit integers), horizontal pass
// into a uint8 temporary, then the vertical pass; each output = clip8((2^21 + sum k_i * in_i) >> 22).
// The coefficient tables are built on the HOST with the same double arithmetic (libm sin) when the plan is
// created and uploaded once; the two passes are integer kernels, so the result is bit-identical to Pillow's
// (tests/test_pyramid.py: against oracle/pillow_resize.py and against the installed Pillow).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "../../include/md2_loss.h"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
double lanczos_filter(double x) {
  /* truncated sinc */
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole-image box
struct Coeffs {
  int ksize = 0;
  std::vector<int> bounds;   // out_size x (xmin, count)
  std::vector<int> kk;       // out_size x ksize
};
Coeffs precompute(int in_size, int out_size) {
  Coeffs c;
  double scale, filterscale;
  filterscale = scale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 3.0 * filterscale;
  c.ksize = (int)ceil(support) * 2 + 1;
  c.bounds.assign((size_t)out_size * 2, 0);
  c.kk.assign((size_t)out_size * c.ksize, 0);
  std::vector<double> w(c.ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      w[x] = lanczos_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double k = (ww != 0.0) ? w[x] / ww : w[x];
      c.kk[(size_t)xx * c.ksize + x] = (k < 0) ? (int)(-0.5 + k * (1 << kPrecisionBits)) : (int)(0.5 + k * (1 << kPrecisionBits));
    }
    c.bounds[2 * xx] = xmin;
    c.bounds[2 * xx + 1] = xmax;
  }
  return c;
}

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= kPrecisionBits;
  return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One separable pass.  Image addressing: element (b, c, y, x) at b*bs + c*cs + y*ys + x*xs for both layouts.
// AXIS 0: resample x (in_w -> out_w, rows unchanged); AXIS 1: resample y.
template <int AXIS>
__global__ void md2_resample_u8(const unsigned char* __restrict__ in, unsigned char* __restrict__ out,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                int batch, int in_h, int in_w, int out_h, int out_w, int hwc) {
  const long long n = (long long)batch * out_h * out_w;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)(i % out_w);
  const int y = (int)((i / out_w) % out_h);
  const int b = (int)(i / ((long long)out_w * out_h));
  const long long ics = hwc ? 1 : (long long)in_h * in_w, ixs = hwc ? 3 : 1, iys = (long long)in_w * ixs;
  const long long ocs = hwc ? 1 : (long long)out_h * out_w, oxs = hwc ? 3 : 1, oys = (long long)out_w * oxs;
  const unsigned char* src = in + (long long)b * 3 * in_h * in_w;
  unsigned char* dst = out + (long long)b * 3 * out_h * out_w + y * oys + x * oxs;
  const int o = AXIS == 0 ? x : y;
  const int lo = __ldg(bounds + 2 * o), cnt = __ldg(bounds + 2 * o + 1);
  const int* k = kk + (size_t)o * ksize;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const unsigned char* p = AXIS == 0 ? src + y * iys + lo * ixs : src + lo * iys + x * ixs;
  const long long step = AXIS == 0 ? ixs : iys;
  for (int j = 0; j < cnt; ++j) {
    const int w = __ldg(k + j);
    s0 += (int)__ldg(p) * w;
    s1 += (int)__ldg(p + ics) * w;
    s2 += (int)__ldg(p + 2 * ics) * w;
    p += step;
  }
  dst[0] = clip8(s0);
  dst[ocs] = clip8(s1);
  dst[2 * ocs] = clip8(s2);
}

}  // namespace

struct md2_resize_plan {
  int in_h, in_w, out_h, out_w;
  int ksize_x, ksize_y;
  int *bounds_x, *kk_x, *bounds_y, *kk_y;   // device
};

extern "C" {

int md2_resize_plan_create(int in_h, int in_w, int out_h, int out_w, md2_resize_plan** plan) {
  if (!plan || in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1) return MD2_ERR_INVALID_ARGUMENT;
  md2_resize_plan* p = (md2_resize_plan*)calloc(1, sizeof(md2_resize_plan));
  if (!p) return MD2_ERR_INVALID_ARGUMENT;
  p->in_h = in_h; p->in_w = in_w; p->out_h = out_h; p->out_w = out_w;
  const Coeffs cx = precompute(in_w, out_w), cy = precompute(in_h, out_h);
  p->ksize_x = cx.ksize; p->ksize_y = cy.ksize;
  auto up = [](const std::vector<int>& v, int** d) {
    if (cudaMalloc((void**)d, v.size() * sizeof(int)) != cudaSuccess) return false;
    return cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  if (!up(cx.bounds, &p->bounds_x) || !up(cx.kk, &p->kk_x) || !up(cy.bounds, &p->bounds_y) || !up(cy.kk, &p->kk_y)) {
    md2_resize_plan_destroy(p);
    return MD2_ERR_CUDA;
  }
  *plan = p;
  return MD2_OK;
}

void md2_resize_plan_destroy(md2_resize_plan* p) {
  if (!p) return;
  cudaFree(p->bounds_x); cudaFree(p->kk_x); cudaFree(p->bounds_y); cudaFree(p->kk_y);
  free(p);
}

int md2_resize_scratch_bytes(const md2_resize_plan* p, int batch, size_t* bytes) {
  if (!p || !bytes || batch < 1) return MD2_ERR_INVALID_ARGUMENT;
  *bytes = (size_t)batch * 3 * p->in_h * p->out_w;     // the horizontal pass's uint8 temporary
  return MD2_OK;
}

int md2_resize_lanczos_u8(const md2_resize_plan* p, const unsigned char* in, unsigned char* out, void* scratch,
                          size_t scratch_bytes, int batch, int hwc, void* stream) {
  if (!p || !in || !out || batch < 1) return MD2_ERR_INVALID_ARGUMENT;
  cudaStream_t s = (cudaStream_t)stream;
  const bool need_x = p->in_w != p->out_w, need_y = p->in_h != p->out_h;
  if (need_x && need_y) {
    size_t need = 0;
    md2_resize_scratch_bytes(p, batch, &need);
    if (!scratch || scratch_bytes < need) return MD2_ERR_WORKSPACE_TOO_SMALL;
  }
  const unsigned char* cur = in;
  int cur_w = p->in_w;
  if (need_x) {
    unsigned char* dst = need_y ? (unsigned char*)scratch : out;
    const long long n = (long long)batch * p->in_h * p->out_w;
    md2_resample_u8<0><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cur, dst, p->bounds_x, p->kk_x, p->ksize_x, batch,
                                                                    p->in_h, p->in_w, p->in_h, p->out_w, hwc);
    cur = dst;
    cur_w = p->out_w;
  }
  if (need_y) {
    const long long n = (long long)batch * p->out_h * p->out_w;
    md2_resample_u8<1><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cur, out, p->bounds_y, p->kk_y, p->ksize_y, batch,
                                                                    p->in_h, cur_w, p->out_h, p->out_w, hwc);
  }
  if (!need_x && !need_y) {
    if (cudaMemcpyAsync(out, in, (size_t)batch * 3 * p->in_h * p->in_w, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return MD2_ERR_CUDA;
  }
  return cudaGetLastError() == cudaSuccess ? MD2_OK : MD2_ERR_CUDA;
}

}  // extern "C"
