// md2_emu.cpp - lock-step HOST emulator of the marching kernels.  TEST INFRASTRUCTURE ONLY.
//
// Compiles monodepth2_b200/csrc/md2_core.cuh (the per-lane arithmetic the CUDA kernels
// run) with g++ and drives it the way md2_kernels.cu does: 32 lanes per warp job in an
// array, neighbour exchange by array index instead of warp shuffle.  It lets the
// tile/halo/reflection/adjoint logic be compared with the oracle in a container that has
// no GPU.  Nothing under monodepth2_b200/ links or loads this file; it is never a
// fallback for the CUDA library.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../monodepth2_b200/csrc/md2_core.cuh"
#include "../../monodepth2_b200/csrc/md2_plan.h"

using namespace md2;

template <int NSRC, bool NOSSIM>
static void emu_identity_t(const Params& P) {
  for (int b = 0; b < P.B; ++b)
    for (int seg = 0; seg < P.nseg_id; ++seg)
      for (int band = 0; band < P.nband_id; ++band) {
        const int y0 = seg * P.id_rows, y1 = std::min(y0 + P.id_rows, P.H);
        IdLane<NSRC> L[32];
        for (int l = 0; l < 32; ++l) { id_init(L[l], P, band * kIdCols, l); id_prefetch(L[l], P, b, y0 - 1); }
        for (int t = y0 - 1; t <= y1; ++t) {
          for (int l = 0; l < 32; ++l) id_stage_a(L[l], P, b, t, l, y0, y1);
          IdXchg<NSRC> X[34];
          memset(X, 0, sizeof(X));
          for (int l = 0; l < 32; ++l) {
            memcpy(X[l + 1].pr, L[l].pr, sizeof(L[l].pr));
            memcpy(X[l + 1].tg, L[l].tg, sizeof(L[l].tg));
          }
          for (int l = 0; l < 32; ++l) id_stage_b<NSRC, NOSSIM>(L[l], P, b, t, l, y0, y1, X[l], X[l + 2]);
        }
      }
}

template <int NSRC>
static void emu_identity(const Params& P) {
  if (P.no_ssim) emu_identity_t<NSRC, true>(P); else emu_identity_t<NSRC, false>(P);
}

template <class C>
static void emu_march(const Params& P) {
  std::vector<F4> ring((size_t)32 * (kRing * C::STASH4 + 2 * C::NB4));
  for (int s = 0; s < P.S; ++s)
    for (int b = 0; b < P.B; ++b)
      for (int seg = 0; seg < P.nseg; ++seg)
        for (int band = 0; band < P.nband; ++band) {
          const int jy0 = seg * P.seg_rows;
          const WarpJob J = make_job(P, s, b, band * kOwnCols, jy0, std::min(jy0 + P.seg_rows, P.H));
          std::fill(ring.begin(), ring.end(), make_f4(0.f, 0.f, 0.f, 0.f));
          Lane<C> L[32];
          Stash st[32];
          for (int l = 0; l < 32; ++l) {
            lane_init(L[l], P, J, l);
            st[l].base = ring.data() + l;
            st[l].bring = ring.data() + 32 * kRing * C::STASH4 + l;
            st[l].stride = 32;
            st[l].tbar = nullptr; st[l].t0 = 0;
            bring_reset<C>(st[l]);
          }
          auto run_c = [&](int t) {
            Xchg2<C> X2[34];
            memset(X2, 0, sizeof(X2));
            X2[0].tag = X2[33].tag = -1;
            for (int l = 0; l < 32; ++l) {
              memcpy(X2[l + 1].coef, L[l].coef, sizeof(L[l].coef));
              X2[l + 1].tag = L[l].tag;
            }
            for (int l = 0; l < 32; ++l) stage_c(L[l], P, J, t, l, X2[l], X2[l + 2], st[l]);
          };
          // same software-pipelined order as md2_march
          for (int t = J.y0 - 2; t <= J.y1 + 1; ++t) {
            for (int l = 0; l < 32; ++l) stage_a_issue(L[l], P, J, t);
            if (C::GRAD && t > J.y0 - 2) run_c(t - 1);
            for (int l = 0; l < 32; ++l) stage_a_finish(L[l], P, J, t, st[l]);
            Xchg1<C> X1[34];
            memset(X1, 0, sizeof(X1));
            for (int l = 0; l < 32; ++l) {
              memcpy(X1[l + 1].pr, L[l].pr, sizeof(L[l].pr));
              memcpy(X1[l + 1].tg, L[l].tg, sizeof(L[l].tg));
            }
            for (int l = 0; l < 32; ++l) stage_b(L[l], P, J, t, l, X1[l], X1[l + 2]);
          }
          if (C::GRAD) run_c(J.y1 + 1);
          float ls = 0.f;
          for (int l = 0; l < 32; ++l) ls += L[l].loss;
          P.acc[acc_photo(s)] += (double)ls;
          if (C::GRAD)
            for (int f = 0; f < C::NSRC; ++f) {
              if (!P.pose_grad[f]) continue;
              float sum[12] = {0};
              for (int l = 0; l < 32; ++l) {
                float dP[12];
                lane_dP(L[l], P, J, f, dP);
                for (int k = 0; k < 12; ++k) sum[k] += dP[k];
              }
              for (int k = 0; k < 12; ++k) P.acc[acc_dP(P, J.ps, b, f, k)] += (double)sum[k];
            }
        }
}

template <int NSRC, bool NOSSIM>
static void emu_march_k(const Params& P) {
  const int key = (P.avg ? 4 : 0) | (P.automask ? 2 : 0) | (P.want_grad ? 1 : 0);
  switch (key) {
    case 0: emu_march<Cfg<NSRC, false, false, false, NOSSIM>>(P); break;
    case 1: emu_march<Cfg<NSRC, false, false, true, NOSSIM>>(P); break;
    case 2: emu_march<Cfg<NSRC, false, true, false, NOSSIM>>(P); break;
    case 3: emu_march<Cfg<NSRC, false, true, true, NOSSIM>>(P); break;
    case 4: emu_march<Cfg<NSRC, true, false, false, NOSSIM>>(P); break;
    case 5: emu_march<Cfg<NSRC, true, false, true, NOSSIM>>(P); break;
    case 6: emu_march<Cfg<NSRC, true, true, false, NOSSIM>>(P); break;
    default: emu_march<Cfg<NSRC, true, true, true, NOSSIM>>(P); break;
  }
}
template <int NSRC>
static void emu_march_n(const Params& P) {
  if (P.no_ssim) emu_march_k<NSRC, true>(P); else emu_march_k<NSRC, false>(P);
}

// decision export for the decision-locked fp64 test (protocol P4): pass NULL to switch it off
extern "C" void md2_emu_set_debug(md2::DebugSink* sink) { md2::debug_sink() = sink; }

extern "C" int md2_emu_workspace_bytes(const md2_problem* p, size_t* bytes) {
  const int st = validate(p);
  if (st != MD2_OK) return st;
  *bytes = make_layout(p).total;
  return MD2_OK;
}

// Same contract as md2_view_synthesis_loss, but every pointer is HOST memory.
extern "C" int md2_emu_view_synthesis_loss(const md2_problem* p, const md2_tensors* t, void* workspace,
                                           size_t workspace_bytes) {
  int st = validate(p);
  if (st != MD2_OK) return st;
  if (workspace_bytes < make_layout(p).total) return MD2_ERR_WORKSPACE_TOO_SMALL;
  Params P;
  st = fill_params(p, t, workspace, &P);
  if (st != MD2_OK) return st;
  // 1. prologue
  for (int i = 0; i < acc_count(P); ++i) P.acc[i] = 0.0;
  if (P.posecnn) {
    for (int s = 0; s < P.S; ++s)
      for (int b = 0; b < P.B; ++b) {
        const int Hs = P.H >> P.lvl[s], Ws = P.W >> P.lvl[s];
        double a = 0.0;
        for (int i = 0; i < P.H * P.W; ++i) a += upsample_at(P.disp[s] + (size_t)b * Hs * Ws, P.lvl[s], Hs, Ws, i / P.W, i % P.W);
        P.acc[acc_updisp(P, s, b)] = a;
        posecnn_mid(P, s, b);
      }
  }
  for (int ps = 0; ps < P.npose; ++ps)
    for (int b = 0; b < P.B; ++b)
      for (int f = 0; f < P.nsrc; ++f) setup_projection(P, ps, b, f);
  if (P.pmask_on)
    for (int s = 0; s < P.S; ++s)
      for (int b = 0; b < P.B; ++b)
        for (int f = 0; f < P.nsrc; ++f)
          for (int i = 0; i < P.H * P.W; ++i) P.acc[acc_bce(P, s)] += (double)pmask_up_pixel(P, s, b, f, i / P.W, i % P.W);
  if (!P.automask)   // with automasking the identity pass writes the RGBx texels itself
    for (int img = 0; img <= P.nsrc; ++img)
      for (int b = 0; b < P.B; ++b)
        for (int p2 = 0; p2 < P.H * P.W; ++p2) pack_pixel(P, img, b, p2);
  // 2. disparity means
  for (int s = 0; s < P.S; ++s)
    for (int b = 0; b < P.B; ++b) {
      const int n = (P.H >> P.lvl[s]) * (P.W >> P.lvl[s]);
      double a = 0.0;
      for (int i = 0; i < n; ++i) a += P.disp[s][(size_t)b * n + i];
      P.acc[acc_dispsum(P, s, b)] = a;
    }
  // 3. identity
  if (P.automask) {
    if (P.nsrc == 1) emu_identity<1>(P); else if (P.nsrc == 2) emu_identity<2>(P); else if (P.nsrc == 3) emu_identity<3>(P);
    else emu_identity<4>(P);
  }
  // 4. smoothness
  for (int s = 0; s < P.S; ++s)
    for (int b = 0; b < P.B; ++b) {
      const int Hs = P.H >> P.lvl[s], Ws = P.W >> P.lvl[s], n = Hs * Ws;
      const float m = (float)(P.acc[acc_dispsum(P, s, b)] / (double)n) + 1e-7f;
      const float inv_m = 1.0f / m;
      for (int i = 0; i < n; ++i) {
        float e0, e1, g;
        smooth_pixel(P, s, b, i / Ws, i % Ws, inv_m, e0, e1, g);
        P.gn[s][(size_t)b * n + i] = g;
        P.acc[acc_smx(P, s, b)] += e0;
        P.acc[acc_smy(P, s, b)] += e1;
        P.acc[acc_dot(P, s, b)] += (double)g * P.disp[s][(size_t)b * n + i];
      }
    }
  for (int s = 0; s < P.S; ++s)
    for (int b = 0; b < P.B; ++b) smooth_scalars(P, s, b, P.smsc[2 * (s * P.B + b)], P.smsc[2 * (s * P.B + b) + 1]);
  // 5. march
  if (P.nsrc == 1) emu_march_n<1>(P); else if (P.nsrc == 2) emu_march_n<2>(P); else if (P.nsrc == 3) emu_march_n<3>(P);
  else emu_march_n<4>(P);
  // 6. final
  final_scalars(P);
  if (P.want_grad && P.posecnn)
    for (int b = 0; b < P.B; ++b) {
      final_pose_posecnn(P, b);
      if (P.lvl[0] == 0)   // level 0 is finished by the marching pass; the other levels add the constant in the adjoint
        for (int i = 0; i < P.H * P.W; ++i) P.grad_disp[0][(size_t)b * P.H * P.W + i] += P.gmidc[b];
    }
  if (P.want_grad && P.pmask_on)
    for (int s = 0; s < P.S; ++s) {
      if (!P.grad_pmask[s]) continue;
      const int Hs = P.H >> P.lvl[s], Ws = P.W >> P.lvl[s];
      for (int b = 0; b < P.B; ++b)
        for (int f = 0; f < P.nsrc; ++f)
          for (int i = 0; i < Hs * Ws; ++i) pmask_grad_pixel(P, s, b, f, i / Ws, i % Ws);
    }
  if (P.want_grad) {
    for (int b = 0; b < P.B; ++b)
      for (int f = 0; f < P.nsrc; ++f) final_grad_T(P, b, f);
    for (int s = 0; s < P.S; ++s)
      for (int b = 0; b < P.B; ++b) {
        const int lv = P.lvl[s], Hs = P.H >> lv, Ws = P.W >> lv, n = Hs * Ws;
        const float inv_m2 = P.smsc[2 * (s * P.B + b)], dterm = P.smsc[2 * (s * P.B + b) + 1];
        for (int i = 0; i < n; ++i) {
          float up = 0.f;
          const int Y = i / Ws, X = i % Ws;
          if (lv == 0) continue;   // finished by the marching pass
          else if (lv == 1) { for (int j = 0; j < 2; ++j) up += upsample_adjoint_part<2>(P, s, b, Y, X, j); }
          else if (lv == 2) { for (int j = 0; j < 4; ++j) up += upsample_adjoint_part<4>(P, s, b, Y, X, j); }
          else { for (int j = 0; j < 8; ++j) up += upsample_adjoint_part<8>(P, s, b, Y, X, j); }
          P.grad_disp[s][(size_t)b * n + i] = up + final_smooth_grad(P, s, b, i, inv_m2, dterm);
        }
      }
  }
  return MD2_OK;
}
