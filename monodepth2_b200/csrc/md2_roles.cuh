// md2_roles.cuh - role-specialised form of the marching kernel (the default since round 2).
//
// md2_march (md2_kernels.cu) keeps the whole fused forward + adjoint state of a pixel column in one
// thread: ~240 registers, 8 warps per SM, and every warp runs one long dependent instruction stream
// per image row (measured in round 1: a warp alone needs ~3 250 cycles per row for 1 183
// instructions; the kernel is bound by that latency, not by issue slots or bandwidth).
//
// Here the same per-lane arithmetic (md2_core.cuh / md2_pack2.cuh, unchanged) is split over warps that
// work on the SAME band of 32 columns, a few image rows apart, and hand rows to each other through
// shared memory:
//   role A  stage_a_issue + stage_a_finish of row t      (disparity -> depth -> projection -> gather ->
//           interpolation); publishes the row (target, pred, d pred / d (ix,iy), u, v, z) in a ring.
//           Rows are independent in this stage, so it is dealt to kRoleAWarps warps round-robin; each of
//           them issues the gather of its next row, crosses the CTA barrier, and only then interpolates:
//           the gather latency is covered by the other warps' work.
//   role B  stage_b of row t-1   (window sums, SSIM + L1, per-pixel minimum / automask, loss, the
//           SSIM-adjoint coefficients of the winner); reads its own and its neighbours' row from the
//           ring (no shuffles), publishes coefficients + winner in a second ring; its streamed inputs
//           (identity loss, tie-break noise) are loaded two rows ahead
//   role C  stage_c of row t-2   (3x3 box adjoint, d loss / d pred, grid-sample and projection
//           adjoints, d loss / d disparity, pose sums)
// One CTA = one band, one bar.sync per image row.  Each warp carries only its own role's rolling
// state (126-168 registers instead of 240-255, no spills with 1-3 sources), so 16 warps are resident per
// SM instead of 8, and the dependent chains of a row run concurrently on different warps instead of back
// to back in one.
#pragma once

#include "md2_core.cuh"
#include "md2_pack2.cuh"

namespace md2 {

// Measured on B200 (mono 640x192 x 12, md2_march_roles alone, profiles/r02_optimization_log.md):
//   1 A warp, 5 CTAs per SM (126 registers)  0.366 ms      2 A warps, 4 CTAs (128 registers)  0.402 ms
//   1 A warp, 4 CTAs (150 registers)         0.430 ms      3 A warps, 3 CTAs                  0.498 ms
//   1 A warp + cp.async gather, 5 CTAs       0.388 ms      round-1 kernel (one warp per band) 0.419 ms
#ifndef MD2_ROLE_MIN_CTAS
#define MD2_ROLE_MIN_CTAS 5
#endif
#ifndef MD2_ROLE_A_WARPS
#define MD2_ROLE_A_WARPS 1
#endif
#ifndef MD2_ROLE_ASYNC_TAPS
#define MD2_ROLE_ASYNC_TAPS 0
#endif
constexpr int kRoleAWarps = MD2_ROLE_A_WARPS;
// who issues the TMA bulk copy of the target row (kernels with gradients): 1 = role C, right after its last read of
// the ring slot the row goes to, one period before role A fills the rest of that slot (role A is the role every
// barrier waits for, role C has ~50 % slack: profiles/r02f_march_roles.txt); 0 = role A, in the period of the row
#ifndef MD2_ROLE_TMA_IN_C
#define MD2_ROLE_TMA_IN_C 1
#endif
// who turns the disparity of a row into depth (packed two-source kernels with gradients): 1 = role C, one period ahead
// of role A, handing the row of depths over in shared memory (two 32-float rows); role A then runs no disparity
// loads, no up-sampling and no reciprocal.  Same instruction sequence on the same values: results are bit-identical.
// Measured (gpurun_out/t_times.log): role A 328 -> 213 instructions per row and no longer the role the barrier waits
// for, but role C (403 -> 470) now is, and ptxas puts its next-row disparity loads in front of the consumer of the
// previous ones, whose register the address computation reuses: 0.475 ms against 0.366 ms.  Off.
#ifndef MD2_ROLE_Z_IN_C
#define MD2_ROLE_Z_IN_C 0
#endif
#if defined(MD2_ROLE_PACKED_C) || defined(MD2_ROLE_A_PIPE) || (MD2_ROLE_ASYNC_TAPS != 0)     // (experimental role forms keep their own depth path)
#undef MD2_ROLE_Z_IN_C
#define MD2_ROLE_Z_IN_C 0
#endif
// cp.async form of role A: 1 = the disparity taps of row t+2 are put in flight inside the issue phase of row t+1,
// right after those of row t+1 are consumed (a whole period to land); 0 = after the finish phase (round-2 first form)
#ifndef MD2_ASYNC_ROW_STEP
#define MD2_ASYNC_ROW_STEP 1
#endif
#ifndef MD2_ROLE_MIN_CTAS3
#define MD2_ROLE_MIN_CTAS3 4      // three sources, or --avg_reprojection, with gradients: 168 registers (__maxnreg__), no spills, 4 CTAs per SM; three sources AND --avg_reprojection: 3 CTAs (224 registers)
#endif

// the role kernel keeps the backward box sums of every source count in registers (role C has room)
#ifndef MD2_ROLE_ZUP
#define MD2_ROLE_ZUP 1
#endif
template <class C0>
struct RoleOf : C0 {
  static constexpr bool BSMEM = false;
  static constexpr bool ZUP = (MD2_ROLE_ZUP != 0);      // depth planes from md2_depth_up (see Cfg::ZUP)
  // branch-free stages B / C also under --avg_reprojection (measured at 640x192 x 12: 0.442 -> 0.413 ms; the two-source
  // per-pixel-minimum kernels lose with it: mono 0.326 -> 0.339 ms, --disable_automasking 0.358 -> 0.375 ms)
#ifdef MD2_ROLE_STRAIGHT
  static constexpr bool STRAIGHT = true;
#else
  static constexpr bool STRAIGHT = C0::STRAIGHT || C0::AVG;
#endif
};
// loops of the packed roles without a branch around the row body (the first / last periods, in which a role only
// crosses the barrier, are peeled): with the branch ptxas may park the wait for the loads a role keeps in flight across
// the barrier on that branch, at the top of the loop, instead of at their first use (measured: 29 % of role B's time)
#ifndef MD2_ROLE_PEEL
#define MD2_ROLE_PEEL 1
#endif
// (one instantiation - 3 sources, --disable_automasking, SSIM, gradients - spills 12 bytes at its 168-register cap in
// the peeled form and keeps the branchy loops)
template <class C>
__host__ __device__ constexpr bool role_peel() {
  return (MD2_ROLE_PEEL != 0) && !(C::NSRC == 3 && !C::AUTOMASK && C::GRAD && !C::AVG && !C::NOSSIM);
}
// role A of the packed kernel computes the projection of row t+1 in the shadow of row t's gather (see role_a2)
#ifndef MD2_ROLE_A_AHEAD
#define MD2_ROLE_A_AHEAD 2
#endif
// the packed two-source role kernel runs its roles over PairedOf<C>: ring layout in the register pairs role B reads
// (see Cfg::PAIRED; the experimental packed role C and the free-running kernel keep the per-source layout)
#ifndef MD2_ROLE_PAIRED
#define MD2_ROLE_PAIRED 1
#endif
#if defined(MD2_ROLE_PACKED_C) || defined(MD2_WITH_FLOW)
#undef MD2_ROLE_PAIRED
#define MD2_ROLE_PAIRED 0
#endif
template <class C0>
struct PairedOf : C0 {
  static constexpr bool PAIRED = (MD2_ROLE_PAIRED != 0);
};

template <class C>
struct RoleCfg {
  static constexpr int NA = kRoleAWarps;                          // warps sharing role A
  static constexpr int NROLES = NA + (C::GRAD ? 2 : 1);
  static constexpr int THREADS = 32 * NROLES;
  static constexpr int RING = C::GRAD ? 5 : 2;                   // rows in flight: A writes t, B reads t-1, C reads t-4
  // register budget per instantiation (no spills anywhere): 4 sources 255, 3 sources or --avg_reprojection with
  // gradients 168 (4 CTAs per SM: the 3-source kernel at 176 registers dropped to 3 CTAs and 0.81 ms instead of
  // 0.53 ms at 640x192 x 12), both together 224, everything else 136
  static constexpr int MIN_CTAS = (C::NSRC >= 4) ? 2 : (C::NSRC >= 3 && C::AVG && C::GRAD) ? 3 : ((C::NSRC >= 3 || C::AVG) && C::GRAD) ? MD2_ROLE_MIN_CTAS3 : MD2_ROLE_MIN_CTAS;
  // registers per thread such that MIN_CTAS CTAs fit an SM (64 K registers, allocation unit 8 per thread); given as
  // __maxnreg__ rather than as the min-blocks argument of __launch_bounds__, under which ptxas picks 168 registers
  // plus a few bytes of spills for the 3-source kernels although 224 would fit
  // NOTE (5 CTAs of 96 threads): the arithmetic says 136, but a kernel that really USES 129-136 registers is given 4 CTAs
  // per SM (ncu launch__occupancy_limit_registers, measured on md2_march_mb: 0.43 instead of 0.33 ms).  Every shipped
  // 5-CTA instantiation uses <= 128 under this cap (tests/test_capi_symbols.py::test_register_budget reads the ptxas
  // log); capping at 128 instead makes ptxas schedule role B 6 instructions longer (0.329 vs 0.324 ms).
  static constexpr int MAXREG = ((65536 / (MIN_CTAS * THREADS)) / 8) * 8 > 255 ? 255 : ((65536 / (MIN_CTAS * THREADS)) / 8) * 8;
  static constexpr int NCF4 = (9 * C::NCS + 1 + 3) / 4;          // coefficient sets + winner tag, 16-byte fields
  static constexpr int STASH_F4 = RING * C::STASH4 * 32;
  static constexpr int COEF_F4 = C::GRAD ? 2 * NCF4 * 32 : 0;
  // NA == 1: role A keeps two rows of bilinear taps in flight through cp.async (LDGSTS) into this buffer
  static constexpr int TAPROW_F4 = C::NSRC * 4 * 32;
  static constexpr bool ASYNC_TAPS = (NA == 1) && (MD2_ROLE_ASYNC_TAPS != 0);
  static constexpr int TAP_F4 = ASYNC_TAPS ? 2 * TAPROW_F4 : 0;
  static constexpr int BAR_F4 = (RING * 8 + 15) / 16;            // one mbarrier per ring slot (TMA-staged target row)
  static constexpr int UNI_F4 = 3;                                // lane-invariant projection rows (role_a2_pipe)
  static constexpr int ZR_F4 = 16;                                // two rows of 32 depths (role C -> role A)
  static constexpr int SMEM_F4 = STASH_F4 + COEF_F4 + TAP_F4 + BAR_F4 + UNI_F4 + ZR_F4;
};

template <int NT>
__device__ __forceinline__ void role_sync() {
  asm volatile("bar.sync 0, %0;" ::"r"(NT) : "memory");
}

// Interior bands (WarpJob::staged): the 32 RGBx target texels of row t are one contiguous 512-byte run of the
// texel array and field 0 of a ring slot is one contiguous 512-byte run of shared memory, so the row is staged by a
// single TMA bulk copy (cp.async.bulk, UBLKCP) issued by one lane of role A; its completion is tracked by the
// slot's mbarrier, which roles B and C wait on before they read the field.  No lane loads or stores the target.
template <class C, class ST>
__device__ __forceinline__ void stage_target_row(const WarpJob& J, const ST& st, int t, int lane) {
  if (J.staged && lane == 0) {
    const int tr = reflect_clamp(t, J.H);
    const int slot = st.slot(t);
    mbar_expect_tx(st.tbar + slot, 32 * 16);
    tma_bulk_g2s(&st.at(slot, 0, C::STASH4), J.tgt4 + 4 * (tr * J.W + J.x0 - 2), 32 * 16, st.tbar + slot);
  }
}
template <class C, bool PACKED>
__host__ __device__ constexpr bool z_in_c() { return PACKED && C::GRAD && !C::ZUP && (MD2_ROLE_A_WARPS == 1) && (MD2_ROLE_Z_IN_C != 0); }
// role C: depth of row t (the arithmetic of stage_a_issue, same order) -> shared memory; puts the disparity taps of
// row t + 1 in flight after those of row t are consumed
template <class C>
__device__ __forceinline__ void c_publish_z(Lane<C>& L, const Params& P, const WarpJob& J, int t, float* zdst) {
  const int tr = reflect_clamp(t, J.H);
  float D;
  if (J.s == 0) {
    D = L.nd[0];
  } else {
    float syr = fmaf(J.rs, (float)tr + 0.5f, -0.5f);
    syr = syr < 0.0f ? 0.0f : syr;
    const float l1 = syr - (float)(int)syr, l0 = 1.0f - l1;
    const float top = up_blend(L.ul0, L.nd[0], L.ul1, L.nd[1]);
    const float bot = up_blend(L.ul0, L.nd[2], L.ul1, L.nd[3]);
    D = up_blend(l0, top, l1, bot);
  }
  prefetch_row<C, false>(L, J, t + 1);
  const float sd = MD2_FADD(P.a_disp, MD2_FMUL(P.c_disp, D));
  *zdst = MD2_RCP(sd);
}
template <class C>
__host__ __device__ constexpr bool tma_in_c() { return C::GRAD && (MD2_ROLE_TMA_IN_C != 0); }
// role C, period i (after its reads of slot(t0 + i - 4) == slot(t0 + i + 1)): stage the target row role A publishes next
template <class C, class ST>
__device__ __forceinline__ void stage_next_target_row(const WarpJob& J, const ST& st, int t, int t1, int lane) {
  if (tma_in_c<C>() && J.staged && t <= t1) {
    __syncwarp();
    stage_target_row<C>(J, st, t, lane);
  }
}
template <class C, class ST>
__device__ __forceinline__ void wait_target_row(const WarpJob& J, const ST& st, int t) {
  if (J.staged && t >= st.t0) mbar_wait(st.tbar + st.slot(t), st.parity(t));     // (rows before t0 are never staged)
}

// ---- role A, warp k of NA: rows t0 + k, t0 + k + NA, ...  Row t must be in the ring before the barrier that
// ends period (t - t0); its gather is issued NA - 1 barriers earlier (NA == 1: in the same period).
// NA == 1: one warp, two rows in flight.  The gather of row t+1 is issued (cp.async into shared memory: completion is
// tracked by the async-group counter, so the wait for row t does not also wait for row t+1, which a register
// destination sharing one scoreboard with it would) before row t is interpolated; the loop is unrolled by two so that
// the two Flight records alternate without moves.
template <class C, class ST>
__device__ __forceinline__ void role_a_async(const Params& P, const WarpJob& J, int lane, const ST& st, F4* tapbuf,
                                             int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane<C> L;
  lane_init(L, P, J, lane);
  Flight<C> F[2];
  stage_a_issue<C, false, 0, true, true>(L, F[0], P, J, t0, tapbuf);
  cp_async_commit();
  prefetch_row<C, false>(L, J, t0 + 1);
  auto half = [&](int t, Flight<C>& Fc, Flight<C>& Fn, F4* bufc, F4* bufn) {
    if (t + 1 <= t1) stage_a_issue<C, false, 0, true, true>(L, Fn, P, J, t + 1, bufn);
    cp_async_commit();
    if (t <= t1) {
      if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
      cp_async_wait<1>();
#pragma unroll
      for (int f = 0; f < C::NSRC; ++f)
#pragma unroll
        for (int q = 0; q < 4; ++q) Fc.tap[f][q] = bufc[(f * 4 + q) * kLanes];
      stage_a_finish<C, ST, true>(L, Fc, P, J, t, st);
      prefetch_row<C, false>(L, J, t + 2);
    }
    role_sync<RC::THREADS>();
  };
#pragma unroll 1
  for (int p = 0; p < nit; p += 2) {
    half(t0 + p, F[0], F[1], tapbuf, tapbuf + RC::TAPROW_F4);
    if (p + 1 < nit) half(t0 + p + 1, F[1], F[0], tapbuf + RC::TAPROW_F4, tapbuf);
  }
  cp_async_wait<0>();
}

template <class C, class ST>
__device__ __forceinline__ void role_a(const Params& P, const WarpJob& J, int lane, int k, const ST& st, int t0, int t1, int nit) {
  constexpr int NA = RoleCfg<C>::NA;
  Lane<C> L;
  lane_init(L, P, J, lane);
  // NA > 1: the disparity taps of a warp's next row are put in flight in its FINISH phase and consumed in its next
  // ISSUE phase, one barrier later (a load issued in the phase that still consumes the previous instance of the
  // same load shares its scoreboard with it: measured as long-scoreboard stalls right after the issue)
  if (NA > 1) {
    prefetch_row<C, false>(L, J, t0 + k);
    if (k < NA - 1) {
      stage_a_issue<C, false, 0, true>(L, P, J, t0 + k);
      prefetch_row<C, false>(L, J, t0 + k + NA);
    }
  }
  if (NA == 1 && role_peel<C>()) {
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
      stage_a_issue<C, false, 1, true>(L, P, J, t);
      stage_a_finish<C, ST, true>(L, P, J, t, st);
      role_sync<RoleCfg<C>::THREADS>();
    }
#pragma unroll 1
    for (int p = t1 - t0 + 1; p < nit; ++p) role_sync<RoleCfg<C>::THREADS>();
    return;
  }
#pragma unroll 1
  for (int p = 0; p < nit; ++p) {
    const int t = t0 + p;
    if (NA == 1) {
      if (t <= t1) {
        if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
        stage_a_issue<C, false, 1, true>(L, P, J, t);
        stage_a_finish<C, ST, true>(L, P, J, t, st);
      }
    } else {
      int ph = (p - k) % NA;
      ph = ph < 0 ? ph + NA : ph;
      if (ph == 0) {
        if (t <= t1) {
          if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
          stage_a_finish<C, ST, true>(L, P, J, t, st);
          if (t >= t0 + NA - 1) prefetch_row<C, false>(L, J, t + NA);    // (the prologue did it for the first rows)
        }
      } else if (ph == 1) {
        if (t + NA - 1 <= t1) stage_a_issue<C, false, 0, true>(L, P, J, t + NA - 1);
      }
    }
    role_sync<RoleCfg<C>::THREADS>();
  }
}

// ---- role B, one row: reads its own and its neighbours' row t from the ring, stage_b, publishes the
// coefficients + winner into the 16-byte fields at `o`
template <class C, class ST>
__device__ __forceinline__ void b_step(Lane<C>& L, const Params& P, const WarpJob& J, int lane, const ST& st, F4* o,
                                       int t, int ol, int orr) {
  typedef RoleCfg<C> RC;
  const int slot = st.slot(t);
  Xchg1<C> lf, rt;
  wait_target_row<C>(J, st, t);
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
#pragma unroll
    for (int k = 0; k < RC::NCF4; ++k) o[k * 32] = make_f4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
  // identity loss / noise of the NEXT row are put in flight after this row's values have been consumed: a load
  // issued before their last use would share its scoreboard with them
  load_identity_row(L, J, t + 1);
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
  load_identity_row(L, J, t0);
  if (role_peel<C>()) {
    role_sync<RC::THREADS>();
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      b_step(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr);
      role_sync<RC::THREADS>();
    }
#pragma unroll 1
    for (int i = t1 - t0 + 2; i < nit; ++i) role_sync<RC::THREADS>();
  } else {
#pragma unroll 1
    for (int i = 0; i < nit; ++i) {
      const int t = t0 + i - 1;
      if (i >= 1 && t <= t1) b_step(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr);
      role_sync<RC::THREADS>();
    }
  }
  const float ls = warp_sum(L.loss);
  if (lane == 0) atomicAdd(&P.acc[acc_photo(J.s)], (double)ls);
}

// ---- role C, one row: reads its own and its neighbours' coefficients of step t from `q`, stage_c
template <class C, class ST>
__device__ __forceinline__ void c_step(Lane<C>& L, const Params& P, const WarpJob& J, int lane, const ST& st, const F4* q,
                                       int t, int ol, int orr) {
  typedef RoleCfg<C> RC;
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
  wait_target_row<C>(J, st, t - 2);
  stage_c(L, P, J, t, lane, lf, rt, st);
}

template <class C>
__device__ __forceinline__ void c_reduce(const Lane<C>& L, const Params& P, const WarpJob& J, int lane) {
#pragma unroll
  for (int f = 0; f < C::NSRC; ++f) {
    if (!P.pose_grad[f]) continue;
    float dP[12];
    lane_dP(L, P, J, f, dP);
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float v = warp_sum(dP[k]);
      if (lane == 0) atomicAdd(&P.acc[acc_dP(P, J.ps, J.b, f, k)], (double)v);
    }
  }
}

// ---- role C: adjoint rows; the coefficients of step t were written one step earlier
template <class C, class ST>
__device__ __forceinline__ void role_c(const Params& P, const WarpJob& J, int lane, const ST& st, const F4* cring,
                                       int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane<C> L;
  lane_init(L, P, J, lane);
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
stage_next_target_row<C>(J, st, t0, t1, lane);
  if (role_peel<C>()) {
    stage_next_target_row<C>(J, st, t0 + 1, t1, lane);
    role_sync<RC::THREADS>();
    stage_next_target_row<C>(J, st, t0 + 2, t1, lane);
    role_sync<RC::THREADS>();
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      c_step(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr);
      stage_next_target_row<C>(J, st, t + 3, t1, lane);
      role_sync<RC::THREADS>();
    }
    c_reduce(L, P, J, lane);
    return;
  }
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i - 2;
    if (i >= 2) c_step(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr);
    stage_next_target_row<C>(J, st, t0 + i + 1, t1, lane);
    role_sync<RC::THREADS>();
  }
  c_reduce(L, P, J, lane);
}

// ------------------------------------------------------------------ packed-fp32 roles (two sources, per-pixel minimum)
// Same three roles over the f32x2 stage functions of md2_pack2.cuh (FFMA2 / FADD2 / FMUL2): the scalar FP32
// instructions issue at one per two cycles per scheduler on this part, and role B is the longest of the three
// (494 instructions per row, 312 of them on the fma pipe); packed over the two sources it comes down to the size
// of the other two.
template <class C, class ST>
__device__ __forceinline__ void role_a2_async(const Params& P, const WarpJob& J, int lane, const ST& st, F4* tapbuf,
                                              int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane2<C> L;
  lane_init2(L, P, J, lane);
  Flight2 F[2];
  stage_a_issue2<C, false, MD2_ASYNC_ROW_STEP, true, true>(L, F[0], P, J, t0, tapbuf);
  cp_async_commit();
  if (MD2_ASYNC_ROW_STEP == 0) prefetch_row2<C, false>(L, J, t0 + 1);
  auto half = [&](int t, Flight2& Fc, Flight2& Fn, F4* bufc, F4* bufn) {
    if (t + 1 <= t1) stage_a_issue2<C, false, MD2_ASYNC_ROW_STEP, true, true>(L, Fn, P, J, t + 1, bufn);
    cp_async_commit();
    if (t <= t1) {
      if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
      cp_async_wait<1>();
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int q = 0; q < 4; ++q) Fc.tap[f][q] = bufc[(f * 4 + q) * kLanes];
      stage_a_finish2<C, ST, true>(L, Fc, P, J, t, st);
      if (MD2_ASYNC_ROW_STEP == 0) prefetch_row2<C, false>(L, J, t + 2);
    }
    role_sync<RC::THREADS>();
  };
#pragma unroll 1
  for (int p = 0; p < nit; p += 2) {
    half(t0 + p, F[0], F[1], tapbuf, tapbuf + RC::TAPROW_F4);
    if (p + 1 < nit) half(t0 + p + 1, F[1], F[0], tapbuf + RC::TAPROW_F4, tapbuf);
  }
  cp_async_wait<0>();
}

template <class C, class ST>
__device__ __forceinline__ void role_a2(const Params& P, const WarpJob& J, int lane, int k, const ST& st, const float* zring,
                                        int t0, int t1, int nit) {
  constexpr int NA = RoleCfg<C>::NA;
  Lane2<C> L;
  lane_init2(L, P, J, lane);
  if (z_in_c<C, true>()) role_sync<RoleCfg<C>::THREADS>();      // role C has published the depths of row t0
  if (NA > 1) {
    prefetch_row2<C, false>(L, J, t0 + k);
    if (k < NA - 1) {
      stage_a_issue2<C, false, 0, true>(L, P, J, t0 + k);
      prefetch_row2<C, false>(L, J, t0 + k + NA);
    }
  }
#if MD2_ROLE_A_AHEAD
  // (--no_ssim: the kernel would land on 136 registers - the 4-CTA trap - or spill under a 128 cap; it keeps the plain form)
  if (NA == 1 && role_peel<C>() && !z_in_c<C, true>() && C::ZUP && tma_in_c<C>() && !C::NOSSIM) {
    // projection one row ahead: the depth -> projection -> bilinear-cell chain of row t+1 (no memory access besides the
    // depth of row t+2 put in flight) runs while the gather of row t is in flight; a period then starts with the gather
#if MD2_ROLE_A_AHEAD == 2
    // two records, two rows per trip: no copy of the record between its computation and its use
    Proj2 Ra, Rb;
    stage_a_proj2<C>(L, Ra, P, J, t0, L.nd[0]);
    prefetch_row2<C, false>(L, J, t0 + 1);
    auto half = [&](int t, Proj2& Rc, Proj2& Rn) {
      stage_a_gather2<C, false>(L, L.fl, Rc, J, t);
      const float zn = L.nd[0];
      prefetch_row2<C, false>(L, J, t + 2);
      stage_a_proj2<C>(L, Rn, P, J, t + 1, zn);
      L.fl.cz = Rc.cz; L.fl.cu = Rc.cu; L.fl.cv = Rc.cv; L.fl.cwx = Rc.cwx; L.fl.cwy = Rc.cwy; L.fl.cgx = Rc.cgx; L.fl.cgy = Rc.cgy;
      stage_a_finish2<C, ST, true>(L, P, J, t, st);
      role_sync<RoleCfg<C>::THREADS>();
    };
#pragma unroll 1
    for (int t = t0; t <= t1; t += 2) {
      half(t, Ra, Rb);
      if (t + 1 <= t1) half(t + 1, Rb, Ra);
    }
#else
    Proj2 R;
    stage_a_proj2<C>(L, R, P, J, t0, L.nd[0]);
    prefetch_row2<C, false>(L, J, t0 + 1);
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      stage_a_gather2<C>(L, L.fl, R, J, t);
      const float zn = L.nd[0];                       // depth of row t+1 (in flight since the previous period)
      prefetch_row2<C, false>(L, J, t + 2);
      stage_a_proj2<C>(L, R, P, J, t + 1, zn);
      stage_a_finish2<C, ST, true>(L, P, J, t, st);
      role_sync<RoleCfg<C>::THREADS>();
    }
#endif
#pragma unroll 1
    for (int p = t1 - t0 + 1; p < nit; ++p) role_sync<RoleCfg<C>::THREADS>();
    return;
  }
#endif
  if (NA == 1 && role_peel<C>() && !z_in_c<C, true>()) {
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
      stage_a_issue2<C, false, 1, true>(L, P, J, t);
      stage_a_finish2<C, ST, true>(L, P, J, t, st);
      role_sync<RoleCfg<C>::THREADS>();
    }
#pragma unroll 1
    for (int p = t1 - t0 + 1; p < nit; ++p) role_sync<RoleCfg<C>::THREADS>();
    return;
  }
#pragma unroll 1
  for (int p = 0; p < nit; ++p) {
    const int t = t0 + p;
    if (NA == 1) {
      if (t <= t1) {
        if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
        if (z_in_c<C, true>()) stage_a_issue2<C, false, 0, true, false, false, true>(L, L.fl, P, J, t, nullptr, nullptr, zring + (p & 1) * 32);
        else stage_a_issue2<C, false, 1, true>(L, P, J, t);
        stage_a_finish2<C, ST, true>(L, P, J, t, st);
      }
    } else {
      int ph = (p - k) % NA;
      ph = ph < 0 ? ph + NA : ph;
      if (ph == 0) {
        if (t <= t1) {
          if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
          stage_a_finish2<C, ST, true>(L, P, J, t, st);
          if (t >= t0 + NA - 1) prefetch_row2<C, false>(L, J, t + NA);
        }
      } else if (ph == 1) {
        if (t + NA - 1 <= t1) stage_a_issue2<C, false, 0, true>(L, P, J, t + NA - 1);
      }
    }
    role_sync<RoleCfg<C>::THREADS>();
  }
}

// NA == 1, register form of the two-rows-in-flight schedule: the gather of row t+1 is issued BEFORE row t is
// interpolated, into the other of two Flight records (loop unrolled by two, so the records alternate without moves and
// the loads of even and odd rows are different static instructions: ptxas gives them different scoreboards, and the
// interpolation of row t does not wait for the gather of row t+1).  Measured reason (profiles/r02n): with issue and
// finish of the SAME row back to back, role A spends 39 % of its time waiting for the gather it just issued, and A is
// the role every barrier waits for (B waits 32 % of its time at the barrier, C 50 %).
template <class C, class ST>
__device__ __forceinline__ void role_a2_pipe(const Params& P, const WarpJob& J, int lane, const ST& st, P2* uni,
                                             int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane2<C> L;
  lane_init2(L, P, J, lane);
  // lane-invariant projection rows -> shared memory (read back as broadcast loads in every issue phase)
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { uni[i] = L.qb[i]; uni[3 + i] = L.p4[i]; }
  }
  __syncwarp();
  Flight2 F[2];
  stage_a_issue2<C, false, 1, true, false, true>(L, F[0], P, J, t0, nullptr, uni);
  auto half = [&](int t, Flight2& Fc, Flight2& Fn) {
    if (t <= t1) {
      if (!tma_in_c<C>()) stage_target_row<C>(J, st, t, lane);
      // border bands (not staged by TMA): the target texel of row t, in flight across the issue phase of row t+1
      F4 tg = make_f4(0.f, 0.f, 0.f, 0.f);
      if (!J.staged) tg = MD2_LDS4(J.tgt4 + 4 * (reflect_clamp(t, J.H) * J.W + L.xi));
      if (t + 1 <= t1) stage_a_issue2<C, false, 1, true, false, true>(L, Fn, P, J, t + 1, nullptr, uni);
      Fc.ctg = tg;
      stage_a_finish2<C, ST, true>(L, Fc, P, J, t, st);
    }
    role_sync<RC::THREADS>();
  };
#pragma unroll 1
  for (int p = 0; p < nit; p += 2) {
    half(t0 + p, F[0], F[1]);
    if (p + 1 < nit) half(t0 + p + 1, F[1], F[0]);
  }
}

// Role B's streamed inputs (identity loss and tie-break noise of the two sources) through four running pointers that
// step one image row per loop trip: load_identity_row2 rebuilds the four addresses from the job description every row
// (~38 instructions for 4 loads in the shipped SASS: ptxas rematerialises the base pointers instead of keeping them)
#ifndef MD2_ROLE_ID_PTRS
#define MD2_ROLE_ID_PTRS 1
#endif
struct IdRows {
  const float* id0; const float* id1; const float* nz0; const float* nz1;
};
template <class C>
__device__ __forceinline__ void id_rows_init(IdRows& R, const Lane2<C>& L, const WarpJob& J, int t) {
  const int yw = t - 1;
  const int pix = (yw < 0 ? 0 : (yw >= J.H ? J.H - 1 : yw)) * J.W + L.xi;
  R.id0 = J.idl + pix; R.id1 = R.id0 + J.plane;
  R.nz0 = J.noise + pix; R.nz1 = R.nz0 + J.plane;
}
// loads the row positioned for step t, then moves to the row of step t + 1 (window row t - 1 clamped to the image)
template <class C>
__device__ __forceinline__ void id_rows_load(Lane2<C>& L, IdRows& R, const WarpJob& J, int t) {
  L.idv[0] = MD2_LDS1(R.id0); L.idv[1] = MD2_LDS1(R.id1);
  L.nzv[0] = MD2_LDS1(R.nz0); L.nzv[1] = MD2_LDS1(R.nz1);
  const int inc = (t >= 1 && t <= J.H - 1) ? J.W : 0;
  R.id0 += inc; R.id1 += inc; R.nz0 += inc; R.nz1 += inc;
#ifdef MD2_ROLE_ID_PREFETCH
  // the rows the NEXT call loads are pulled into L2 now (no destination register, no scoreboard): identity loss and
  // noise are read once per scale and come from HBM; one period is not always enough for the load itself
  asm volatile("prefetch.global.L2 [%0];" ::"l"(R.id0));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(R.id1));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(R.nz0));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(R.nz1));
#endif
}
template <class C>
__host__ __device__ constexpr bool id_ptrs() { return C::AUTOMASK && (MD2_ROLE_ID_PTRS != 0); }

template <class C, class ST>
__device__ __forceinline__ void b_step2(Lane2<C>& L, const Params& P, const WarpJob& J, int lane, const ST& st, F4* o,
                                        int t, int ol, int orr, IdRows* idr = nullptr) {
  const int slot = st.slot(t);
  Xchg1P<C> lf, rt;
  wait_target_row<C>(J, st, t);
  const F4* p0 = &st.at(slot, 0, C::STASH4);
  const F4* p1 = &st.at(slot, 1, C::STASH4);
  const F4* p2_ = &st.at(slot, 4, C::STASH4);
#ifdef MD2_ROLE_SHFL_XCHG
  // neighbours by warp shuffle instead of a second and third LDS.128 of the same ring row: 24 shared-memory
  // wavefronts fewer per row (the LSU pipe is the busiest unit of this kernel, profiles/r02n_march_keys.txt)
  const F4 tc = p0[0], ac = p1[0], bc_ = p2_[0];
  auto up3 = [](const F4& v) { return make_f4(__shfl_up_sync(kFull, v.x, 1), __shfl_up_sync(kFull, v.y, 1), __shfl_up_sync(kFull, v.z, 1), 0.f); };
  auto dn3 = [](const F4& v) { return make_f4(__shfl_down_sync(kFull, v.x, 1), __shfl_down_sync(kFull, v.y, 1), __shfl_down_sync(kFull, v.z, 1), 0.f); };
  const F4 tl = up3(tc), tr = dn3(tc), al = up3(ac), ar = dn3(ac), bl = up3(bc_), br = dn3(bc_);
  (void)ol; (void)orr;
#else
  const F4 tc = p0[0], tl = p0[ol], tr = p0[orr];
  const F4 ac = p1[0], al = p1[ol], ar = p1[orr];          // source 0: pred (r,g,b), u
  const F4 bc_ = p2_[0], bl = p2_[ol], br = p2_[orr];      // source 1
#endif
  L.tgrg = p2(tc.x, tc.y); L.tgb = tc.z;
  lf.tgrg = p2(tl.x, tl.y); lf.tgb = tl.z;
  rt.tgrg = p2(tr.x, tr.y); rt.tgb = tr.z;
  if (C::PAIRED) {      // fields 1 / 4 = (r0, g0, r1, g1) / (b0, b1, u0, u1): the pairs as they are
    L.pr[0] = p2(ac.x, ac.y); L.pr[1] = p2(ac.z, ac.w); L.pr[2] = p2(bc_.x, bc_.y);
    lf.pr[0] = p2(al.x, al.y); lf.pr[1] = p2(al.z, al.w); lf.pr[2] = p2(bl.x, bl.y);
    rt.pr[0] = p2(ar.x, ar.y); rt.pr[1] = p2(ar.z, ar.w); rt.pr[2] = p2(br.x, br.y);
  } else {
    L.pr[0] = p2(ac.x, ac.y); L.pr[1] = p2(bc_.x, bc_.y); L.pr[2] = p2(ac.z, bc_.z);
    lf.pr[0] = p2(al.x, al.y); lf.pr[1] = p2(bl.x, bl.y); lf.pr[2] = p2(al.z, bl.z);
    rt.pr[0] = p2(ar.x, ar.y); rt.pr[1] = p2(br.x, br.y); rt.pr[2] = p2(ar.z, br.z);
  }
  // branch-free form (every lane computes its window, results are selected): 0.3257 vs 0.3295 ms with the divergent one
#ifdef MD2_B2_DIVERGENT
  stage_b2(L, P, J, t, lane, lf, rt);
#else
  stage_b2_straight(L, P, J, t, lane, lf, rt);
#endif
  if (C::GRAD) {
    o[0] = make_f4(L.cf[0].x, L.cf[0].y, L.cf[1].x, L.cf[1].y);
    o[32] = make_f4(L.cf[2].x, L.cf[2].y, L.cfb[0], L.cfb[1]);
    o[64] = make_f4(L.cfb[2], __int_as_float(L.tag), 0.f, 0.f);
  }
  if (id_ptrs<C>() && idr) id_rows_load(L, *idr, J, t + 1);
  else load_identity_row2(L, J, t + 1);
}

template <class C, class ST>
__device__ __forceinline__ void role_b2(const Params& P, const WarpJob& J, int lane, const ST& st, F4* cring,
                                        int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane2<C> L;
  lane_init2(L, P, J, lane);
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
  IdRows idr;
  if (id_ptrs<C>()) {
    id_rows_init(idr, L, J, t0);
    id_rows_load(L, idr, J, t0);
  } else load_identity_row2(L, J, t0);
  if (z_in_c<C, true>()) role_sync<RC::THREADS>();
  if (role_peel<C>()) {
    role_sync<RC::THREADS>();                                     // period 0: role A publishes row t0
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      b_step2(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr, id_ptrs<C>() ? &idr : nullptr);
      role_sync<RC::THREADS>();
    }
#pragma unroll 1
    for (int i = t1 - t0 + 2; i < nit; ++i) role_sync<RC::THREADS>();
  } else {
#pragma unroll 1
    for (int i = 0; i < nit; ++i) {
      const int t = t0 + i - 1;
      if (i >= 1 && t <= t1) b_step2(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr, id_ptrs<C>() ? &idr : nullptr);
      role_sync<RC::THREADS>();
    }
  }
  const float ls = warp_sum(L.loss);
  if (lane == 0) atomicAdd(&P.acc[acc_photo(J.s)], (double)ls);
}

template <class C, class ST>
__device__ __forceinline__ void c_step2(Lane2<C>& L, const Params& P, const WarpJob& J, int lane, const ST& st, const F4* q,
                                        int t, int ol, int orr) {
  auto unpack = [](const F4& a, const F4& b, const F4& c, P2* cf, float* cfb, int& tag) {
    cf[0] = p2(a.x, a.y); cf[1] = p2(a.z, a.w); cf[2] = p2(b.x, b.y);
    cfb[0] = b.z; cfb[1] = b.w; cfb[2] = c.x;
    tag = __float_as_int(c.y);
  };
  Xchg2P<C> lf, rt;
  unpack(q[0], q[32], q[64], L.cf, L.cfb, L.tag);
  unpack(q[ol], q[32 + ol], q[64 + ol], lf.cf, lf.cfb, lf.tag);
  unpack(q[orr], q[32 + orr], q[64 + orr], rt.cf, rt.cfb, rt.tag);
  wait_target_row<C>(J, st, t - 2);
  stage_c2(L, P, J, t, lane, lf, rt, st);
}

// scalar stage_c fed from the packed coefficient layout of b_step2 (the packed stage_c2 needs a few registers more
// than the 128 that five CTAs per SM leave: 44 bytes of spills; the scalar one fits)
template <class C, class ST>
__device__ __forceinline__ void c_step_from_packed(Lane<C>& L, const Params& P, const WarpJob& J, int lane, const ST& st,
                                                   const F4* q, int t, int ol, int orr) {
  static_assert(C::NCS == 1, "per-pixel minimum");
  auto unpack = [](const F4& a, const F4& b, const F4& c, float* coef, int& tag) {
    coef[0] = a.x; coef[3] = a.y; coef[1] = a.z; coef[4] = a.w; coef[2] = b.x; coef[5] = b.y;
    coef[6] = b.z; coef[7] = b.w; coef[8] = c.x;
    tag = __float_as_int(c.y);
  };
  Xchg2<C> lf, rt;
  unpack(q[0], q[32], q[64], L.coef[0], L.tag);
#ifdef MD2_ROLE_SHFL_XCHG
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    lf.coef[0][k] = __shfl_up_sync(kFull, L.coef[0][k], 1);
    rt.coef[0][k] = __shfl_down_sync(kFull, L.coef[0][k], 1);
  }
  lf.tag = __shfl_up_sync(kFull, L.tag, 1);
  rt.tag = __shfl_down_sync(kFull, L.tag, 1);
  (void)ol; (void)orr;
#else
  unpack(q[ol], q[32 + ol], q[64 + ol], lf.coef[0], lf.tag);
  unpack(q[orr], q[32 + orr], q[64 + orr], rt.coef[0], rt.tag);
#endif
  wait_target_row<C>(J, st, t - 2);
  stage_c(L, P, J, t, lane, lf, rt, st);
}

template <class C, class ST>
__device__ __forceinline__ void role_c_from_packed(const Params& P, const WarpJob& J, int lane, const ST& st, const F4* cring,
                                                   float* zring, int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane<C> L;
  lane_init(L, P, J, lane);
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
stage_next_target_row<C>(J, st, t0, t1, lane);
  if (z_in_c<C, true>()) {
    c_publish_z(L, P, J, t0, zring);          // (lane_init put the disparity taps of row t0 in flight)
    role_sync<RC::THREADS>();
  }
  if (role_peel<C>() && !z_in_c<C, true>()) {
    // periods 0 and 1: roles A / B fill the rings; this role only stages the next target rows (nit >= 3 always)
    stage_next_target_row<C>(J, st, t0 + 1, t1, lane);
    role_sync<RC::THREADS>();
    stage_next_target_row<C>(J, st, t0 + 2, t1, lane);
    role_sync<RC::THREADS>();
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {                               // period i = t - t0 + 2
      c_step_from_packed(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr);
      stage_next_target_row<C>(J, st, t + 3, t1, lane);
      role_sync<RC::THREADS>();
    }
    c_reduce(L, P, J, lane);
    return;
  }
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i - 2;
    if (i >= 2) c_step_from_packed(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr);
    stage_next_target_row<C>(J, st, t0 + i + 1, t1, lane);
    if (z_in_c<C, true>() && t0 + i + 1 <= t1) c_publish_z(L, P, J, t0 + i + 1, zring + ((i + 1) & 1) * 32);
    role_sync<RC::THREADS>();
  }
  c_reduce(L, P, J, lane);
}

template <class C>
__device__ __forceinline__ void c_reduce2(const Lane2<C>& L, const Params& P, const WarpJob& J, int lane) {
#pragma unroll
  for (int f = 0; f < 2; ++f) {
    if (!P.pose_grad[f]) continue;
    float dP[12];
    lane_dP2(L, P, J, f, dP);
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float v = warp_sum(dP[k]);
      if (lane == 0) atomicAdd(&P.acc[acc_dP(P, J.ps, J.b, f, k)], (double)v);
    }
  }
}

template <class C, class ST>
__device__ __forceinline__ void role_c2(const Params& P, const WarpJob& J, int lane, const ST& st, const F4* cring,
                                        int t0, int t1, int nit) {
  typedef RoleCfg<C> RC;
  Lane2<C> L;
  lane_init2(L, P, J, lane);
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
stage_next_target_row<C>(J, st, t0, t1, lane);
#pragma unroll 1
  for (int i = 0; i < nit; ++i) {
    const int t = t0 + i - 2;
    if (i >= 2) c_step2(L, P, J, lane, st, cring + (t & 1) * RC::NCF4 * 32, t, ol, orr);
    stage_next_target_row<C>(J, st, t0 + i + 1, t1, lane);
    role_sync<RC::THREADS>();
  }
  c_reduce2(L, P, J, lane);
}

#ifdef MD2_ROLE_USE_MINBLOCKS
#define MD2_ROLE_BOUNDS(C) __launch_bounds__(RoleCfg<C>::THREADS, RoleCfg<C>::MIN_CTAS)
#else
#define MD2_ROLE_BOUNDS(C) __launch_bounds__(RoleCfg<C>::THREADS) __maxnreg__(RoleCfg<C>::MAXREG)
#endif
template <class C, bool PACKED>
__global__ void MD2_ROLE_BOUNDS(C) md2_march_roles(Params P) {
  static_assert(!PACKED || (C::NSRC == 2 && !C::AVG), "packed form: two sources, per-pixel minimum");
  typedef RoleCfg<C> RC;
  extern __shared__ float4 smem[];
  const int lane = threadIdx.x & 31;
  // (measured: rotating the warp -> role map from CTA to CTA changes nothing; the hardware does not pin a CTA's
  // warp i to scheduler i % 4 in a way that would pile one role onto one scheduler: per-scheduler issue counts of
  // one launch are within +-10 %)
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
  WarpJob J = make_job(P, js, jb, (r2 - js * P.nband) * kOwnCols, jy0, min(jy0 + P.seg_rows, P.H));
  // bands whose 32 lanes all lie inside the image: the target row is staged by TMA (others read reflected columns)
  // (1-2 sources: measured on one box 0.389 vs 0.419 ms at 640x192, 1.017 vs 1.071 ms at 1024x320; the 3-source
  // kernels, at their register limit, lose: 0.533 vs 0.499 ms, and keep the per-lane loads)
#ifndef MD2_ROLE_NO_TMA
  if (C::NSRC <= 2) J.staged = (J.x0 - 2 >= 0) && (J.x0 + 30 <= J.W);
#endif

  StashT<RC::RING> st;
  st.base = smem + lane;
  st.bring = nullptr;
  st.stride = 32;
  st.tbar = reinterpret_cast<unsigned long long*>(smem + RC::STASH_F4 + RC::COEF_F4 + RC::TAP_F4);
  F4* cring = smem + RC::STASH_F4 + lane;
  F4* tapbuf = smem + RC::STASH_F4 + RC::COEF_F4 + lane;
  float* zring = reinterpret_cast<float*>(smem + RC::STASH_F4 + RC::COEF_F4 + RC::TAP_F4 + RC::BAR_F4 + RC::UNI_F4) + lane;
  const int t0 = J.y0 - 2, t1 = J.y1 + 1;
  st.t0 = t0;
  const int nit = (t1 - t0 + 1) + (C::GRAD ? 2 : 1);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < RC::RING; ++i) mbar_init(st.tbar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if constexpr (RC::ASYNC_TAPS) {
    if (role == 0) {
      if constexpr (PACKED) role_a2_async<C>(P, J, lane, st, tapbuf, t0, t1, nit);
      else role_a_async<C>(P, J, lane, st, tapbuf, t0, t1, nit);
      return;
    }
  }
  if constexpr (PACKED) {
    typedef PairedOf<C> CP;
#ifdef MD2_ROLE_A_PIPE
    if (role < RC::NA) role_a2_pipe<C>(P, J, lane, st, reinterpret_cast<P2*>(smem + RC::STASH_F4 + RC::COEF_F4 + RC::TAP_F4 + RC::BAR_F4), t0, t1, nit);
#else
    if (role < RC::NA) role_a2<CP>(P, J, lane, role, st, zring, t0, t1, nit);
#endif
    else if (role == RC::NA) role_b2<CP>(P, J, lane, st, cring, t0, t1, nit);
#ifdef MD2_ROLE_PACKED_C
    else if (C::GRAD) role_c2<C>(P, J, lane, st, cring, t0, t1, nit);
#else
    else if (C::GRAD) role_c_from_packed<CP>(P, J, lane, st, cring, zring, t0, t1, nit);
#endif
  } else {
    if (role < RC::NA) role_a<C>(P, J, lane, role, st, t0, t1, nit);
    else if (role == RC::NA) role_b<C>(P, J, lane, st, cring, t0, t1, nit);
    else if (C::GRAD) role_c<C>(P, J, lane, st, cring, t0, t1, nit);
  }
}

#ifdef MD2_WITH_MB
// ------------------------------------------------------------------ decoupled roles (mbarrier pipeline)
// md2_march_roles runs its three roles in lock step: one bar.sync per image row, so every period lasts as long as the
// slowest role of THAT row.  Measured on the shipped kernel: roles A and B each wait 12 % of their time at the barrier
// (C 31 %) - neither is the slowest every row, the period is E[max] instead of max E.  Here the same roles (same stage
// functions, same rings, packed two-source kernels with gradients) hand rows over through mbarriers, the way a TMA
// pipeline does - a "full" and an "empty" barrier per ring slot, phase = use count of the slot - and a role waits only
// for the row it needs:
//   A(t)  waits ringE[slot(t)] (role C is done with pixel row t - RING), publishes row t, arrives on ringF[slot(t)]
//   B(t)  waits ringF[slot(t)] (+ the TMA barrier of the target row), waits coefE[cslot(t)], publishes the coefficients
//         of row t, arrives on coefF[cslot(t)]
//   C(t)  waits coefF[cslot(t)], runs the adjoint of pixel row t - 2, arrives on coefE[cslot(t)] and ringE[slot(t - 2)],
//         stages the target row t + 3 by TMA (its slot was released by C(t + 3 - RING + 2) at the latest: RING >= 5)
// Every barrier counts the 32 lanes of the arriving warp (each lane's arrive releases its own shared-memory writes);
// waits are mbarrier.try_wait (the warp sleeps in hardware, no polling of shared memory, no fence).
// Measured (profiles/r02_optimization_log.md): 0.339 ms with 8 ring / 4 coefficient slots against 0.330 ms in lock step, 0.392 ms
// with 5 / 2 or 6 / 2 slots - the barrier waits of the lock-step kernel are not a synchronisation artefact, the roles
// share issue slots and the LSU pipe and slow each other down whichever way they wait.  Compiled only with
// -DMD2_WITH_MB (then the default for the packed two-source kernels with gradients; MD2_MARCH=lockstep switches back).
#ifndef MD2_MB_RING
#define MD2_MB_RING 8
#endif
#ifndef MD2_MB_CRING
#define MD2_MB_CRING 4
#endif
template <class C>
struct MbCfg {
  static constexpr int THREADS = 96;
  static constexpr int RING = MD2_MB_RING;                       // power of two: slot(t) is a mask
  static constexpr int CRING = MD2_MB_CRING;
  static constexpr int MIN_CTAS = MD2_ROLE_MIN_CTAS;
  // 128, not the 136 that 65536 / (5 x 96) suggests: a kernel that really uses 129-136 registers gets 4 CTAs per SM
  // (ncu launch__occupancy_limit_registers = 4; measured 0.43 instead of 0.33 ms)
  static constexpr int MAXREG = RoleCfg<C>::MAXREG > 128 && RoleCfg<C>::MAXREG < 144 ? 128 : RoleCfg<C>::MAXREG;
  static constexpr int NCF4 = RoleCfg<C>::NCF4;
  static constexpr int STASH_F4 = RING * C::STASH4 * 32;
  static constexpr int COEF_F4 = CRING * NCF4 * 32;
  static constexpr int NBAR = 3 * RING + 2 * CRING;              // TMA, ring full, ring empty | coefficients full, empty
  static constexpr int BAR_F4 = (NBAR * 8 + 15) / 16;
  static constexpr int SMEM_F4 = STASH_F4 + COEF_F4 + BAR_F4;
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

template <class C0>
__global__ void __launch_bounds__(MbCfg<PairedOf<C0>>::THREADS) __maxnreg__(MbCfg<PairedOf<C0>>::MAXREG) md2_march_mb(Params P) {
  typedef PairedOf<C0> C;
  typedef MbCfg<C> MC;
  static_assert(C::NSRC == 2 && !C::AVG && C::GRAD, "packed two-source kernels with gradients");
  constexpr int R = MC::RING, CR = MC::CRING;
  extern __shared__ float4 smem[];
  const int lane = threadIdx.x & 31;
  const int role = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
  const int job = blockIdx.x;
  const int per_seg = P.S * P.nband;
  const int per_b = P.nseg * per_seg;
  const int jb = P.B - 1 - job / per_b;
  const int r = job - (job / per_b) * per_b;
  const int seg = r / per_seg;
  const int r2 = r - seg * per_seg;
  const int js = r2 / P.nband;
  const int jy0 = seg * P.seg_rows;
  WarpJob J = make_job(P, js, jb, (r2 - js * P.nband) * kOwnCols, jy0, min(jy0 + P.seg_rows, P.H));
#ifndef MD2_ROLE_NO_TMA
  J.staged = (J.x0 - 2 >= 0) && (J.x0 + 30 <= J.W);
#endif
  StashT<R> st;
  st.base = smem + lane;
  st.bring = nullptr;
  st.stride = 32;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + MC::STASH_F4 + MC::COEF_F4);
  st.tbar = bars;
  unsigned long long* ringF = bars + R;
  unsigned long long* ringE = bars + 2 * R;
  unsigned long long* coefF = bars + 3 * R;
  unsigned long long* coefE = bars + 3 * R + CR;
  F4* cring = smem + MC::STASH_F4 + lane;
  const int t0 = J.y0 - 2, t1 = J.y1 + 1;
  st.t0 = t0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < R; ++i) { mbar_init(bars + i, 1); mbar_init(ringF + i, 32); mbar_init(ringE + i, 32); }
#pragma unroll
    for (int i = 0; i < CR; ++i) { mbar_init(coefF + i, 32); mbar_init(coefE + i, 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;

  if (role == 0) {
    Lane2<C> L;
    lane_init2(L, P, J, lane);
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      const int n = (t - t0) / R;                                           // use count of this row's slot
      stage_a_issue2<C, false, 1, true>(L, P, J, t);
      if (n > 0) mbar_wait(ringE + st.slot(t), (unsigned)((n - 1) & 1));    // (after the gather is in flight)
      stage_a_finish2<C, decltype(st), true>(L, P, J, t, st);
      mbar_arrive(ringF + st.slot(t));
    }
  } else if (role == 1) {
    Lane2<C> L;
    lane_init2(L, P, J, lane);
    IdRows idr;
    if (id_ptrs<C>()) { id_rows_init(idr, L, J, t0); id_rows_load(L, idr, J, t0); }
    else load_identity_row2(L, J, t0);
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      const int i = t - t0, cs = i & (CR - 1), m = i / CR;
      if (m > 0) mbar_wait(coefE + cs, (unsigned)((m - 1) & 1));
      mbar_wait(ringF + st.slot(t), st.parity(t));
      b_step2(L, P, J, lane, st, cring + cs * MC::NCF4 * 32, t, ol, orr, id_ptrs<C>() ? &idr : nullptr);
      mbar_arrive(coefF + cs);
    }
    const float ls = warp_sum(L.loss);
    if (lane == 0) atomicAdd(&P.acc[acc_photo(J.s)], (double)ls);
  } else {
    Lane<C> L;
    lane_init(L, P, J, lane);
    if (J.staged) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (t0 + k <= t1) stage_target_row<C>(J, st, t0 + k, lane);
    }
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
      const int i = t - t0, cs = i & (CR - 1), m = i / CR;
      mbar_wait(coefF + cs, (unsigned)(m & 1));
      c_step_from_packed(L, P, J, lane, st, cring + cs * MC::NCF4 * 32, t, ol, orr);
      mbar_arrive(coefE + cs);
      if (t - 2 >= t0) mbar_arrive(ringE + st.slot(t - 2));
      if (J.staged && t + 3 <= t1) {
        __syncwarp();
        stage_target_row<C>(J, st, t + 3, lane);
      }
    }
    c_reduce(L, P, J, lane);
  }
}

#endif  // MD2_WITH_MB

// ------------------------------------------------------------------ free-running roles
// md2_march_roles above runs its warps in lock step (one bar.sync per image row): every warp waits for the
// slowest role of every row, measured at ~35 % of every warp's time.  Here the same roles are decoupled: the rings
// are deeper, and a warp waits only for the row it needs (a progress counter per role in shared memory, written by
// lane 0 after a warp-level fence, polled by lane 0 of the consumer) or for the ring slot it wants to overwrite.
//   A_k  rows t0 + k, t0 + k + NA, ...   issue + finish back to back (its gather latency is covered by the other
//        warps); before writing ring slot(t): role C must be done with row t - RING (C reads row tC - 2)
//   B    rows t0 .. t1 in order          needs row t from A_(t mod NA); coefficient slot(t): C done with t - CRING
//   C    rows t0 .. t1 in order          needs the coefficients of row t from B (row t - 2 of the ring is then there)
#ifndef MD2_FLOW_RING
#define MD2_FLOW_RING 6
#endif
#ifndef MD2_FLOW_CRING
#define MD2_FLOW_CRING 4
#endif
template <class C>
struct FlowCfg {
  static constexpr int NA = kRoleAWarps;
  static constexpr int NROLES = NA + (C::GRAD ? 2 : 1);
  static constexpr int THREADS = 32 * NROLES;
  static constexpr int RING = MD2_FLOW_RING;
  static constexpr int CRING = MD2_FLOW_CRING;
  static constexpr int MIN_CTAS = RoleCfg<C>::MIN_CTAS;
  static constexpr int NCF4 = RoleCfg<C>::NCF4;
  static constexpr int STASH_F4 = RING * C::STASH4 * 32;
  static constexpr int COEF_F4 = C::GRAD ? CRING * NCF4 * 32 : 0;
  static constexpr int SMEM_F4 = STASH_F4 + COEF_F4 + 2;         // + 8 progress counters
};

struct Flow {
  volatile int* prog;       // [0 .. NA-1]: last row role A_k published, [4]: role B, [5]: role C
  __device__ __forceinline__ void wait_ge(int which, int v, int lane) const {
    if (lane == 0) {
      while (prog[which] < v) { }
      __threadfence_block();
    }
    __syncwarp();
  }
  __device__ __forceinline__ void publish(int which, int v, int lane) const {
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      prog[which] = v;
    }
  }
};

template <class C, bool PACKED>
__global__ void __launch_bounds__(FlowCfg<C>::THREADS, FlowCfg<C>::MIN_CTAS) md2_march_flow(Params P) {
  static_assert(!PACKED || (C::NSRC == 2 && !C::AVG), "packed form: two sources, per-pixel minimum");
  typedef FlowCfg<C> FC;
  constexpr int NA = FC::NA;
  extern __shared__ float4 smem[];
  const int lane = threadIdx.x & 31;
  const int role = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
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

  StashT<FC::RING> st;
  st.base = smem + lane;
  st.bring = nullptr;
  st.stride = 32;
  st.tbar = nullptr;
  st.t0 = J.y0 - 2;
  F4* cring = smem + FC::STASH_F4 + lane;
  Flow fl;
  fl.prog = reinterpret_cast<volatile int*>(smem + FC::STASH_F4 + FC::COEF_F4);
  const int t0 = J.y0 - 2, t1 = J.y1 + 1;
  if (threadIdx.x < 8) fl.prog[threadIdx.x] = t0 - 1;
  __syncthreads();
  const int ol = (lane > 0) ? -1 : 0, orr = (lane < 31) ? 1 : 0;
  // what the writer of ring slot(t) waits for: the last reader of row t - RING
  constexpr int kLast = C::GRAD ? 5 : 4;               // counter of the last role of the chain
  constexpr int kLag = C::GRAD ? 2 : 0;                // that role reads ring row (its t) - kLag

  if (role < NA) {
    const int k = role;
    if constexpr (PACKED) {
      Lane2<C> L;
      lane_init2(L, P, J, lane);
      prefetch_row2<C, false>(L, J, t0 + k);
#pragma unroll 1
      for (int t = t0 + k; t <= t1; t += NA) {
        stage_a_issue2<C, false, NA, true>(L, P, J, t);
        fl.wait_ge(kLast, t - FC::RING + kLag, lane);
        stage_a_finish2<C, decltype(st), true>(L, P, J, t, st);
        fl.publish(k, t, lane);
      }
    } else {
      Lane<C> L;
      lane_init(L, P, J, lane);
      prefetch_row<C, false>(L, J, t0 + k);
#pragma unroll 1
      for (int t = t0 + k; t <= t1; t += NA) {
        stage_a_issue<C, false, NA, true>(L, P, J, t);
        fl.wait_ge(kLast, t - FC::RING + kLag, lane);
        stage_a_finish<C, decltype(st), true>(L, P, J, t, st);
        fl.publish(k, t, lane);
      }
    }
  } else if (role == NA) {
    if constexpr (PACKED) {
      Lane2<C> L;
      lane_init2(L, P, J, lane);
      load_identity_row2(L, J, t0);
#pragma unroll 1
      for (int t = t0; t <= t1; ++t) {
        int ka = (t - t0) % NA;
        fl.wait_ge(ka, t, lane);
        if (C::GRAD) fl.wait_ge(5, t - FC::CRING, lane);
        int cs = (t - t0) % FC::CRING;
        b_step2(L, P, J, lane, st, cring + cs * FC::NCF4 * 32, t, ol, orr);
        fl.publish(4, t, lane);
      }
      const float ls = warp_sum(L.loss);
      if (lane == 0) atomicAdd(&P.acc[acc_photo(J.s)], (double)ls);
    } else {
      Lane<C> L;
      lane_init(L, P, J, lane);
      load_identity_row(L, J, t0);
#pragma unroll 1
      for (int t = t0; t <= t1; ++t) {
        int ka = (t - t0) % NA;
        fl.wait_ge(ka, t, lane);
        if (C::GRAD) fl.wait_ge(5, t - FC::CRING, lane);
        int cs = (t - t0) % FC::CRING;
        b_step(L, P, J, lane, st, cring + cs * FC::NCF4 * 32, t, ol, orr);
        fl.publish(4, t, lane);
      }
      const float ls = warp_sum(L.loss);
      if (lane == 0) atomicAdd(&P.acc[acc_photo(J.s)], (double)ls);
    }
  } else if (C::GRAD) {
    if constexpr (PACKED) {
      Lane2<C> L;
      lane_init2(L, P, J, lane);
#pragma unroll 1
      for (int t = t0; t <= t1; ++t) {
        fl.wait_ge(4, t, lane);
        int cs = (t - t0) % FC::CRING;
        c_step2(L, P, J, lane, st, cring + cs * FC::NCF4 * 32, t, ol, orr);
        fl.publish(5, t, lane);
      }
      c_reduce2(L, P, J, lane);
    } else {
      Lane<C> L;
      lane_init(L, P, J, lane);
#pragma unroll 1
      for (int t = t0; t <= t1; ++t) {
        fl.wait_ge(4, t, lane);
        int cs = (t - t0) % FC::CRING;
        c_step(L, P, J, lane, st, cring + cs * FC::NCF4 * 32, t, ol, orr);
        fl.publish(5, t, lane);
      }
      c_reduce(L, P, J, lane);
    }
  }
}

}  // namespace md2
