"""Drives the CUDA library (through monodepth2_b200.fused_loss, i.e. the C ABI) on a Golden fixture."""
import torch

from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss


def run_cuda(g, rows_per_segment=0, align_corners=False, want_grad=True, side_all=True, dev="cuda:0",
             pose_leaves=False):
    plan = LossPlan(g.B, g.H, g.W, g.frame_ids, avg_reprojection=g.avg_reprojection,
                    disable_automasking=g.disable_automasking, align_corners=align_corners,
                    rows_per_segment=rows_per_segment, no_ssim=g.no_ssim, v1_multiscale=g.v1_multiscale,
                    posecnn=g.posecnn, predictive_mask=g.predictive_mask, scales=g.scales)
    inputs = {k: v.to(dev) for k, v in g.inputs().items()}
    outs, leaves = {}, {}
    for s in g.scales:
        d = g.t("disp__%d" % s).to(dev).requires_grad_(want_grad)
        outs[("disp", s)] = d
        leaves[("disp", s)] = d
    for f in g.frame_ids[1:]:
        if f == "s":
            continue
        if pose_leaves:
            # what PoseDecoder emits: (B,2,1,3), only [:, 0] is used (trainer.py:289-295); no cam_T_cam given,
            # the fused call builds it from the leaves
            aa = torch.zeros(g.B, 2, 1, 3)
            tr = torch.zeros(g.B, 2, 1, 3)
            aa[:, 0] = g.t("axisangle__%s" % f).reshape(g.B, 1, 3)
            tr[:, 0] = g.t("translation__%s" % f).reshape(g.B, 1, 3)
            aa = aa.to(dev).requires_grad_(want_grad)
            tr = tr.to(dev).requires_grad_(want_grad)
            outs[("axisangle", 0, f)], outs[("translation", 0, f)] = aa, tr
            leaves[("axisangle", f)], leaves[("translation", f)] = aa, tr
            continue
        T = g.t("cam_T_cam__%s" % f).to(dev).requires_grad_(want_grad)
        outs[("cam_T_cam", 0, f)] = T
        leaves[("T", f)] = T
        if g.posecnn:          # the leaves are the pose-net outputs (B,n,1,3), trainer.py:366-375
            aa = g.t("axisangle__%s" % f).to(dev).reshape(-1, 1, 1, 3).requires_grad_(want_grad)
            tr = g.t("translation__%s" % f).to(dev).reshape(-1, 1, 1, 3).requires_grad_(want_grad)
            outs[("axisangle", 0, f)], outs[("translation", 0, f)] = aa, tr
            leaves[("axisangle", f)], leaves[("translation", f)] = aa, tr
    if g.predictive_mask:      # the mask decoder's outputs (trainer.py:251-252)
        outs["predictive_mask"] = {}
        for s in g.scales:
            m = g.t("mask__%d" % s).to(dev).requires_grad_(want_grad)
            outs["predictive_mask"][("disp", s)] = m
            leaves[("mask", s)] = m
    noise = [n.to(dev) for n in g.noise()] if g.n_id > 0 else None
    side = None
    if side_all:
        side = {"depth_scales": list(g.scales), "color_scales": list(g.scales), "mask_scales": list(g.scales),
                "grad_updisp_scales": list(g.scales)}
    if want_grad:
        losses = view_synthesis_loss(plan, inputs, outs, noise, side)
        losses["loss"].backward()
    else:
        with torch.no_grad():
            losses = view_synthesis_loss(plan, inputs, outs, noise, side)
    torch.cuda.synchronize()
    return dict(losses=losses, outs=outs, leaves=leaves, side=side)
