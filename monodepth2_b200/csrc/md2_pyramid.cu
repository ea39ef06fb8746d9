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

// ------------------------------------------------------------------ colour augmentation
// torchvision ColorJitter on PIL images as MonoDataset applies it (/root/reference/datasets/mono_dataset.py:60-70,
// 136,169-176): adjust_brightness / contrast / saturation = Pillow's Image.blend(degenerate, image, factor)
// (libImaging/Blend.c: in1 + alpha * (in2 - in1) in C float, truncated), the "L" conversion of Convert.c, and
// adjust_hue = RGB -> HSV -> h += shift (mod 256) -> RGB with Convert.c's float / double mix.  Every operation is
// written with explicit _rn intrinsics (no contraction) so that the bytes equal Pillow's; oracle/color_jitter.py is the
// restatement the tests hold this to (itself pinned against the installed Pillow, the HSV pair on all 2^24 colours).
struct Rgb8 { int r, g, b; };

__device__ __forceinline__ int cj_to_l(const Rgb8& c) { return (c.r * 19595 + c.g * 38470 + c.b * 7471 + 0x8000) >> 16; }

__device__ __forceinline__ int cj_blend1(int in1, int in2, float a, bool interp) {
  const float t = __fadd_rn((float)in1, __fmul_rn(a, (float)(in2 - in1)));
  if (interp) return (int)t;
  return t <= 0.0f ? 0 : (t >= 255.0f ? 255 : (int)t);
}
__device__ __forceinline__ Rgb8 cj_blend(const Rgb8& d, const Rgb8& c, float a) {
  const bool interp = a >= 0.0f && a <= 1.0f;
  Rgb8 o;
  o.r = cj_blend1(d.r, c.r, a, interp); o.g = cj_blend1(d.g, c.g, a, interp); o.b = cj_blend1(d.b, c.b, a, interp);
  return o;
}

__device__ __forceinline__ Rgb8 cj_hue(const Rgb8& c, int shift) {
  // rgb2hsv_row
  const int maxc = max(c.r, max(c.g, c.b)), minc = min(c.r, min(c.g, c.b));
  int uh = 0, us = 0;
  const int uv = maxc;
  if (minc != maxc) {
    const float cr = (float)(maxc - minc);
    const float sf = __fdiv_rn(cr, (float)maxc);
    const float rc = __fdiv_rn((float)(maxc - c.r), cr), gc = __fdiv_rn((float)(maxc - c.g), cr), bc = __fdiv_rn((float)(maxc - c.b), cr);
    float h;
    if (c.r == maxc) h = __fsub_rn(bc, gc);
    else if (c.g == maxc) h = __double2float_rn(__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc));
    else h = __double2float_rn(__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc));
    h = __double2float_rn(fmod(__dadd_rn(__ddiv_rn((double)h, 6.0), 1.0), 1.0));
    uh = min(max((int)__dmul_rn((double)h, 255.0), 0), 255);
    us = min(max((int)__dmul_rn((double)sf, 255.0), 0), 255);
  }
  uh = (uh + shift) & 255;
  // hsv2rgb
  Rgb8 o;
  if (us == 0) { o.r = o.g = o.b = uv; return o; }
  const double hf = __ddiv_rn(__dmul_rn((double)(float)uh, 6.0), 255.0);
  const int i = (int)floor(hf);
  const double f = (double)__double2float_rn(__dsub_rn(hf, (double)(float)i));
  const double fs = (double)__double2float_rn(__ddiv_rn((double)(float)us, 255.0));
  const double vf = (double)(float)uv;
  const int p = min(max((int)round(__dmul_rn(vf, __dsub_rn(1.0, fs))), 0), 255);
  const int q = min(max((int)round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, f)))), 0), 255);
  const int t = min(max((int)round(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f))))), 0), 255);
  switch (i % 6) {
    case 0: o.r = uv; o.g = t; o.b = p; break;
    case 1: o.r = q; o.g = uv; o.b = p; break;
    case 2: o.r = p; o.g = uv; o.b = t; break;
    case 3: o.r = p; o.g = q; o.b = uv; break;
    case 4: o.r = t; o.g = p; o.b = uv; break;
    default: o.r = uv; o.g = p; o.b = q; break;
  }
  return o;
}

// applies operations [from, to) of the image's order; `mean` is the contrast gray level (used when op 1 is in range)
__device__ __forceinline__ Rgb8 cj_apply(Rgb8 c, const md2_color_jitter& J, int from, int to, int mean) {
  for (int k = from; k < to; ++k) {
    const int op = J.order[k];
    if (op == 0) { const Rgb8 z = {0, 0, 0}; c = cj_blend(z, c, J.brightness); }
    else if (op == 1) { const Rgb8 m = {mean, mean, mean}; c = cj_blend(m, c, J.contrast); }
    else if (op == 2) { const int l = cj_to_l(c); const Rgb8 g = {l, l, l}; c = cj_blend(g, c, J.saturation); }
    else if (op == 3) c = cj_hue(c, J.hue_shift);
  }
  return c;
}
__device__ __forceinline__ int cj_contrast_pos(const md2_color_jitter& J) {
  for (int k = 0; k < 4; ++k) if (J.order[k] == 1) return k;
  return -1;
}
__device__ __forceinline__ Rgb8 cj_load(const unsigned char* in, long long n, int p, int plane, int hwc) {
  Rgb8 c;
  if (hwc) { const unsigned char* q = in + (n * plane + p) * 3; c.r = __ldg(q); c.g = __ldg(q + 1); c.b = __ldg(q + 2); }
  else { const unsigned char* q = in + n * 3 * plane + p; c.r = __ldg(q); c.g = __ldg(q + plane); c.b = __ldg(q + 2 * (long long)plane); }
  return c;
}

// pass 1 (only for images whose order contains contrast): sum of the "L" image after the operations in front of it
__global__ void __launch_bounds__(256) md2_cj_lsum(const unsigned char* in, const md2_color_jitter* params, unsigned long long* sums,
                                                    int plane, int hwc) {
  const int n = blockIdx.y;
  const md2_color_jitter J = params[n];
  const int cp = cj_contrast_pos(J);
  if (cp < 0) return;
  unsigned int acc = 0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += gridDim.x * blockDim.x)
    acc += (unsigned)cj_to_l(cj_apply(cj_load(in, n, p, plane, hwc), J, 0, cp, 0));
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sums + n, (unsigned long long)acc);
}

// pass 2: the whole chain; mean = int(sum / count + 0.5) in double, as ImageStat.Stat(...).mean does in Python
__global__ void __launch_bounds__(256) md2_cj_apply(const unsigned char* in, unsigned char* out, const md2_color_jitter* params,
                                                     const unsigned long long* sums, int plane, int hwc) {
  const int n = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= plane) return;
  const md2_color_jitter J = params[n];
  const int mean = (int)(__dadd_rn(__ddiv_rn((double)sums[n], (double)plane), 0.5));
  const Rgb8 c = cj_apply(cj_load(in, n, p, plane, hwc), J, 0, 4, mean);
  if (hwc) { unsigned char* q = out + ((long long)n * plane + p) * 3; q[0] = (unsigned char)c.r; q[1] = (unsigned char)c.g; q[2] = (unsigned char)c.b; }
  else { unsigned char* q = out + (long long)n * 3 * plane + p; q[0] = (unsigned char)c.r; q[plane] = (unsigned char)c.g; q[2 * (long long)plane] = (unsigned char)c.b; }
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

int md2_color_jitter_u8(const unsigned char* in, unsigned char* out, const md2_color_jitter* params, void* scratch,
                        size_t scratch_bytes, int n_images, int height, int width, int hwc, void* stream) {
  if (!in || !out || !params || n_images < 1 || height < 1 || width < 1) return MD2_ERR_INVALID_ARGUMENT;
  if (n_images > 65535) return MD2_ERR_UNSUPPORTED;
  if (!scratch || scratch_bytes < (size_t)n_images * sizeof(unsigned long long)) return MD2_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t s = (cudaStream_t)stream;
  const int plane = height * width;
  unsigned long long* sums = (unsigned long long*)scratch;
  if (cudaMemsetAsync(sums, 0, (size_t)n_images * sizeof(unsigned long long), s) != cudaSuccess) return MD2_ERR_CUDA;
  const int bx = (plane + 255) / 256;
  md2_cj_lsum<<<dim3(bx < 64 ? bx : 64, n_images), 256, 0, s>>>(in, params, sums, plane, hwc);
  md2_cj_apply<<<dim3(bx, n_images), 256, 0, s>>>(in, out, params, sums, plane, hwc);
  return cudaGetLastError() == cudaSuccess ? MD2_OK : MD2_ERR_CUDA;
}

}  // extern "C"
