/*
 * md2_debug.h - entry points of the DEBUG build of the library only (libmd2loss_dbg.so, built with
 * -DMD2_DBG_DEVICE -DMD2_BOUNDS_CHECK by `python -m monodepth2_b200.build --debug`).  Test infrastructure:
 * the product library (libmd2loss.so) does not export them and compiles the hooks to nothing.
 *
 * md2_debug_set_sink: the kernels write the discrete decisions they take (bilinear cell + clip masks, per-pixel
 *   winner, SSIM clamp-live bits, L1 / smoothness signs) into the device arrays of `sink` (layout: struct
 *   md2::DebugSink, monodepth2_b200/csrc/md2_core.cuh; all pointers are DEVICE pointers); NULL switches it off.
 *   Used by the decision-locked fp64 test (protocol P4) on the real kernels.
 * md2_debug_oob_count: number of global-memory indices of the marching path that fell outside their tensor
 *   since the last reset (every load / store index is checked in this build).
 */
#ifndef MD2_DEBUG_H_
#define MD2_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif
int md2_debug_set_sink(const void *sink);
int md2_debug_oob_count(unsigned long long *count, int reset);
#ifdef __cplusplus
}
#endif
#endif
