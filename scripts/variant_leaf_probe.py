"""Per-leaf gradient errors of the --predictive_mask variant test case, with the number of elements that differ from the
fp64 oracle by more than 1e-4 of the leaf's maximum (a flip shows up as a handful of large differences)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import test_gpu_variants as T
from helpers import rel_l2
from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
DEV = "cuda:0"
for name, B, H, W, fids, flags in [c for c in T.CASES if c[0] == (sys.argv[1] if len(sys.argv) > 1 else "pmask")]:
    seed = 300 + len(name)
    o_losses, o_leaves, nz = T._oracle(B, H, W, fids, flags, seed, torch.float32)
    d_losses, d_leaves, _ = T._oracle(B, H, W, fids, flags, seed, torch.float64)
    inputs, outs, leaves, _ = T._leaves(B, H, W, fids, flags, seed, torch.float32)
    plan = LossPlan(B, H, W, fids, **flags)
    c_leaves = {k: v.detach().to(DEV).requires_grad_(True) for k, v in leaves.items()}
    c_outs = {k: c_leaves[k] for k in leaves if k[0] == "disp"}
    for f in fids[1:]:
        if f != "s":
            c_outs[("axisangle", 0, f)], c_outs[("translation", 0, f)] = c_leaves[("axisangle", f)], c_leaves[("translation", f)]
    if flags.get("predictive_mask"):
        c_outs["predictive_mask"] = {("disp", s): c_leaves[("mask", s)] for s in range(4)}
    losses = view_synthesis_loss(plan, {k: v.to(DEV) for k, v in inputs.items()}, c_outs, [n.to(DEV) for n in nz] if nz else None)
    losses["loss"].backward()
    print(os.environ.get("MD2_LIB_PATH", "product"), name, "loss", float(losses["loss"]), float(o_losses["loss"]))
    for k in c_leaves:
        g, t, o = c_leaves[k].grad.cpu().double(), d_leaves[k].grad.double(), o_leaves[k].grad.double()
        nbad = int(((g - t).abs() > 1e-4 * t.abs().max()).sum()); nbad_o = int(((o - t).abs() > 1e-4 * t.abs().max()).sum())
        print("   ", k, "mine-vs-f64 %.2e (%d elems off)  ref32-vs-f64 %.2e (%d off)  numel %d" % (rel_l2(g, t), nbad, rel_l2(o, t), nbad_o, g.numel()))
