// md2_roles.cuh - role-specialised form of the marching kernel (the default since round 2).
//
// md2_march (md2_kernels.cu) keeps the whole fused forward + adjoint state of a pixel column in one
// thread: ~240 registers, 8 warps per SM, and every warp runs one long dependent instruction stream
// per image row (measured in round 1: a warp alone needs ~3 250 cycles per row for 1 183
// instructions; the kernel is bound by that latency, not by issue slots or bandwidth).
//
// Here the same per-lane arithmetic (md2_core.cuh, unchanged) is split over three warps that work on
// the SAME band of 32 columns, one image row apart, and hand rows to each other through shared
// memory:
//   role A  stage_a_issue + stage_a_finish of row t      (disparity -> depth -> projection -> gather ->
//           interpolation); publishes the row (target, pred, d pred / d (ix,iy), u, v, z) in a ring
//   role B  stage_b of row t-1   (window sums, SSIM + L1, per-pixel minimum / automask, loss, the
//           SSIM-adjoint coefficients of the winner); reads its own and its neighbours' row from the
//           ring (no shuffles), publishes coefficients + winner in a second ring
//   role C  stage_c of row t-2   (3x3 box adjoint, d loss / d pred, grid-sample and projection
//           adjoints, d loss / d disparity, pose sums)
// One CTA = one band = 3 warps (2 without gradients), one bar.sync per image row.  Each warp carries
// only its own role's rolling state, so the kernel fits 4-5 CTAs (12-15 warps) per SM, and the three
// dependent chains of a row run concurrently on different warps instead of back to back in one.
#pragma once

#include "md2_core.cuh"
#include "md2_pack2.cuh"

namespace md2 {

#ifndef MD2_ROLE_MIN_CTAS
#define MD2_ROLE_MIN_CTAS 4
#endif

// the role kernel keeps the backward box sums of every source count in registers (role C has room)
template <class C0>
struct RoleOf : C0 {
  static constexpr bool BSMEM = false;
};

template <class C>
struct RoleCfg {
  static constexpr int NROLES = C::GRAD ? 3 : 2;
  static constexpr int THREADS = 32 * NROLES;
  static constexpr int RING = C::GRAD ? 5 : 2;                   // rows in flight: A writes t, B reads t-1, C reads t-4
  static constexpr int NCF4 = (9 * C::NCS + 1 + 3) / 4;          // coefficient sets + winner tag, 16-byte fields
  static constexpr int STASH_F4 = RING * C::STASH4 * 32;
  static constexpr int SMEM_F4 = STASH_F4 + (C::GRAD ? 2 * NCF4 * 32 : 0);
};

template <int NT>
__device__ __forceinline__ void role_sync() {
  asm volatile("bar.sync 0, %0;" ::"r"(NT) : "memory");
}

// ---- role A: rows t0 .. t1
template <class C, class ST>
__device__ __forceinline__ void role_a(const Params& P, const WarpJob& J, int lane, const ST& st, int t0, int t1, int nit) {
  Lane<C> L;
  lane_init(L, P, J, lane);
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i;
    if (t <= t1) {
      stage_a_issue<C, false>(L, P, J, t);
      stage_a_finish<C, ST, true>(L, P, J, t, st);
    }
    role_sync<RoleCfg<C>::THREADS>();
  }
}

// ---- role B: window rows; row t of the ring was written one step earlier
template <class C, class ST>
__device__ __forceinline__ void role_b(const Params& P, const WarpJob& J, int lane, const ST& st, F4* cring,
                                       int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane<C> L;
  lane_init(L, P, J, lane);
  // neighbour offsets in the ring (the edge lanes read themselves: their windows are never used)
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i - 1;
    if (i >= 1 && t <= t1) {
      load_identity_row(L, J, t);
      const int slot = st.slot(t);
      Xchg1<C> lf, rt;
      {
        const F4* p = &st.at(slot, 0, C::STASH4);
        const F4 c = p[0], l = p[ol], r = p[orr];
        L.tg[0] = c.x; L.tg[1] = c.y; L.tg[2] = c.z;
        lf.tg[0] = l.x; lf.tg[1] = l.y; lf.tg[2] = l.z;
        rt.tg[0] = r.x; rt.tg[1] = r.y; rt.tg[2] = r.z;
      }
#pragma unroll
      for (int f = 0; f < C::NSRC; ++f) {
        const F4* p = &st.at(slot, 1 + 3 * f, C::STASH4);
        const F4 c = p[0], l = p[ol], r = p[orr];
        L.pr[f][0] = c.x; L.pr[f][1] = c.y; L.pr[f][2] = c.z;
        lf.pr[f][0] = l.x; lf.pr[f][1] = l.y; lf.pr[f][2] = l.z;
        rt.pr[f][0] = r.x; rt.pr[f][1] = r.y; rt.pr[f][2] = r.z;
      }
      stage_b(L, P, J, t, lane, lf, rt);
      if (C::GRAD) {
        float v[4 * RC::NCF4];
#pragma unroll
        for (int k = 0; k < 4 * RC::NCF4; ++k) v[k] = 0.f;
#pragma unroll
        for (int n = 0; n < C::NCS; ++n)
#pragma unroll
          for (int k = 0; k < 9; ++k) v[n * 9 + k] = L.coef[n][k];
        v[9 * C::NCS] = __int_as_float(L.tag);
        F4* o = cring + (t & 1) * RC::NCF4 * 32;
#pragma unroll
        for (int k = 0; k < RC::NCF4; ++k) o[k * 32] = make_f4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      }
    }
    role_sync<RC::THREADS>();
  }
  const float ls = warp_sum(L.loss);
  if (lane == 0) atomicAdd(&P.acc[acc_photo(J.s)], (double)ls);
}

// ---- role C: adjoint rows; the coefficients of step t were written one step earlier
template <class C, class ST>
__device__ __forceinline__ void role_c(const Params& P, const WarpJob& J, int lane, const ST& st, const F4* cring,
                                       int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane<C> L;
  lane_init(L, P, J, lane);
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i - 2;
    if (i >= 2) {
      const F4* q = cring + (t & 1) * RC::NCF4 * 32;
      float vc[4 * RC::NCF4], vl[4 * RC::NCF4], vr[4 * RC::NCF4];
#pragma unroll
      for (int k = 0; k < RC::NCF4; ++k) {
        const F4 c = q[k * 32], l = q[k * 32 + ol], r = q[k * 32 + orr];
        vc[4 * k] = c.x; vc[4 * k + 1] = c.y; vc[4 * k + 2] = c.z; vc[4 * k + 3] = c.w;
        vl[4 * k] = l.x; vl[4 * k + 1] = l.y; vl[4 * k + 2] = l.z; vl[4 * k + 3] = l.w;
        vr[4 * k] = r.x; vr[4 * k + 1] = r.y; vr[4 * k + 2] = r.z; vr[4 * k + 3] = r.w;
      }
      Xchg2<C> lf, rt;
#pragma unroll
      for (int n = 0; n < C::NCS; ++n)
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          L.coef[n][k] = vc[n * 9 + k];
          lf.coef[n][k] = vl[n * 9 + k];
          rt.coef[n][k] = vr[n * 9 + k];
        }
      L.tag = __float_as_int(vc[9 * C::NCS]);
      lf.tag = __float_as_int(vl[9 * C::NCS]);
      rt.tag = __float_as_int(vr[9 * C::NCS]);
      stage_c(L, P, J, t, lane, lf, rt, st);
    }
    role_sync<RC::THREADS>();
  }
#pragma unroll
  for (int f = 0; f < C::NSRC; ++f) {
    if (!P.pose_grad[f]) continue;
    float dP[12];
    lane_dP(L, P, J, f, dP);
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float v = warp_sum(dP[k]);
      if (lane == 0) atomicAdd(&P.acc[acc_dP(P, J.b, f, k)], (double)v);
    }
  }
}

// ------------------------------------------------------------------ packed-fp32 roles (two sources, per-pixel minimum)
// Same three roles over the f32x2 stage functions of md2_pack2.cuh (FFMA2 / FADD2 / FMUL2): the scalar FP32
// instructions issue at one per two cycles per scheduler on this part, and role B is the longest of the three
// (494 instructions per row, 312 of them on the fma pipe); packed over the two sources it comes down to the size
// of the other two.
template <class C, class ST>
__device__ __forceinline__ void role_a2(const Params& P, const WarpJob& J, int lane, const ST& st, int t0, int t1, int nit) {
  Lane2<C> L;
  lane_init2(L, P, J, lane);
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i;
    if (t <= t1) {
      stage_a_issue2<C, false>(L, P, J, t);
      stage_a_finish2<C, ST, true>(L, P, J, t, st);
    }
    role_sync<RoleCfg<C>::THREADS>();
  }
}

template <class C, class ST>
__device__ __forceinline__ void role_b2(const Params& P, const WarpJob& J, int lane, const ST& st, F4* cring,
                                        int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane2<C> L;
  lane_init2(L, P, J, lane);
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i - 1;
    if (i >= 1 && t <= t1) {
      load_identity_row2(L, J, t);
      const int slot = st.slot(t);
      Xchg1P<C> lf, rt;
      const F4* p0 = &st.at(slot, 0, C::STASH4);
      const F4* p1 = &st.at(slot, 1, C::STASH4);
      const F4* p2_ = &st.at(slot, 4, C::STASH4);
      const F4 tc = p0[0], tl = p0[ol], tr = p0[orr];
      const F4 ac = p1[0], al = p1[ol], ar = p1[orr];          // source 0: pred (r,g,b), u
      const F4 bc_ = p2_[0], bl = p2_[ol], br = p2_[orr];      // source 1
      L.tgrg = p2(tc.x, tc.y); L.tgb = tc.z;
      lf.tgrg = p2(tl.x, tl.y); lf.tgb = tl.z;
      rt.tgrg = p2(tr.x, tr.y); rt.tgb = tr.z;
      L.pr[0] = p2(ac.x, ac.y); L.pr[1] = p2(bc_.x, bc_.y); L.pr[2] = p2(ac.z, bc_.z);
      lf.pr[0] = p2(al.x, al.y); lf.pr[1] = p2(bl.x, bl.y); lf.pr[2] = p2(al.z, bl.z);
      rt.pr[0] = p2(ar.x, ar.y); rt.pr[1] = p2(br.x, br.y); rt.pr[2] = p2(ar.z, br.z);
      stage_b2(L, P, J, t, lane, lf, rt);
      if (C::GRAD) {
        F4* o = cring + (t & 1) * RC::NCF4 * 32;
        o[0] = make_f4(L.cf[0].x, L.cf[0].y, L.cf[1].x, L.cf[1].y);
        o[32] = make_f4(L.cf[2].x, L.cf[2].y, L.cfb[0], L.cfb[1]);
        o[64] = make_f4(L.cfb[2], __int_as_float(L.tag), 0.f, 0.f);
      }
    }
    role_sync<RC::THREADS>();
  }
  const float ls = warp_sum(L.loss);
  if (lane == 0) atomicAdd(&P.acc[acc_photo(J.s)], (double)ls);
}

template <class C, class ST>
__device__ __forceinline__ void role_c2(const Params& P, const WarpJob& J, int lane, const ST& st, const F4* cring,
                                        int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane2<C> L;
  lane_init2(L, P, J, lane);
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
  auto unpack = [](const F4& a, const F4& b, const F4& c, P2* cf, float* cfb, int& tag) {
    cf[0] = p2(a.x, a.y); cf[1] = p2(a.z, a.w); cf[2] = p2(b.x, b.y);
    cfb[0] = b.z; cfb[1] = b.w; cfb[2] = c.x;
    tag = __float_as_int(c.y);
  };
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i - 2;
    if (i >= 2) {
      const F4* q = cring + (t & 1) * RC::NCF4 * 32;
      Xchg2P<C> lf, rt;
      unpack(q[0], q[32], q[64], L.cf, L.cfb, L.tag);
      unpack(q[ol], q[32 + ol], q[64 + ol], lf.cf, lf.cfb, lf.tag);
      unpack(q[orr], q[32 + orr], q[64 + orr], rt.cf, rt.cfb, rt.tag);
      stage_c2(L, P, J, t, lane, lf, rt, st);
    }
    role_sync<RC::THREADS>();
  }
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    if (!P.pose_grad[f]) continue;
    float dP[12];
    lane_dP2(L, P, J, f, dP);
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float v = warp_sum(dP[k]);
      if (lane == 0) atomicAdd(&P.acc[acc_dP(P, J.b, f, k)], (double)v);
    }
  }
}

template <class C, bool PACKED>
__global__ void __launch_bounds__(RoleCfg<C>::THREADS, MD2_ROLE_MIN_CTAS) md2_march_roles(Params P) {
  static_assert(!PACKED || (C::NSRC == 2 && !C::AVG), "packed form: two sources, per-pixel minimum");
  typedef RoleCfg<C> RC;
  extern __shared__ float4 smem[];
  const int lane = threadIdx.x & 31;
  const int role = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
  // same job order as md2_march (sample-major, segment, scale, band; last sample first), one job per CTA
  const int job = blockIdx.x;
  const int per_seg = P.S * P.nband;
  const int per_b = P.nseg * per_seg;
  const int jb = P.B - 1 - job / per_b;
  const int r = job - (job / per_b) * per_b;
  const int seg = r / per_seg;
  const int r2 = r - seg * per_seg;
  const int js = r2 / P.nband;
  const int jy0 = seg * P.seg_rows;
  const WarpJob J = make_job(P, js, jb, (r2 - js * P.nband) * kOwnCols, jy0, min(jy0 + P.seg_rows, P.H));

  StashT<RC::RING> st;
  st.base = smem + lane;
  st.bring = nullptr;
  st.stride = 32;
  F4* cring = smem + RC::STASH_F4 + lane;
  const int t0 = J.y0 - 2, t1 = J.y1 + 1;
  const int nit = (t1 - t0 + 1) + (RC::NROLES - 1);
  if constexpr (PACKED) {
    if (role == 0) role_a2<C>(P, J, lane, st, t0, t1, nit);
    else if (role == 1) role_b2<C>(P, J, lane, st, cring, t0, t1, nit);
    else if (C::GRAD) role_c2<C>(P, J, lane, st, cring, t0, t1, nit);
  } else {
    if (role == 0) role_a<C>(P, J, lane, st, t0, t1, nit);
    else if (role == 1) role_b<C>(P, J, lane, st, cring, t0, t1, nit);
    else if (C::GRAD) role_c<C>(P, J, lane, st, cring, t0, t1, nit);
  }
}

}  // namespace md2
