import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import test_gpu_variants as T
from helpers import rel_l2
from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
DEV = "cuda:0"
for name, B, H, W, fids, flags in [("posecnn", 3, 96, 160, [0, -1, 1], dict(posecnn=True)), ("plain", 3, 96, 160, [0, -1, 1], dict()),
                                   ("posecnn_iidseed", 3, 96, 160, [0, -1, 1], dict(posecnn=True))]:
    for seed in ([307] if name != "posecnn_iidseed" else [11, 12, 13]):
        o_losses, o_leaves, nz = T._oracle(B, H, W, fids, flags, seed, torch.float32)
        d_losses, d_leaves, _ = T._oracle(B, H, W, fids, flags, seed, torch.float64)
        inputs, outs, leaves, _ = T._leaves(B, H, W, fids, flags, seed, torch.float32)
        plan = LossPlan(B, H, W, fids, **flags)
        c_leaves = {k: v.detach().to(DEV).requires_grad_(True) for k, v in leaves.items()}
        c_outs = {k: c_leaves[k] for k in leaves if k[0] == "disp"}
        for f in fids[1:]:
            c_outs[("axisangle", 0, f)], c_outs[("translation", 0, f)] = c_leaves[("axisangle", f)], c_leaves[("translation", f)]
        losses = view_synthesis_loss(plan, {k: v.to(DEV) for k, v in inputs.items()}, c_outs, [n.to(DEV) for n in nz])
        losses["loss"].backward()
        print(name, seed, "loss", float(losses["loss"]), float(o_losses["loss"]))
        for k in c_leaves:
            print("   ", k, "mine-vs-f64 %.2e  ref32-vs-f64 %.2e  mine-vs-ref32 %.2e" % (
                rel_l2(c_leaves[k].grad.cpu(), d_leaves[k].grad), rel_l2(o_leaves[k].grad, d_leaves[k].grad),
                rel_l2(c_leaves[k].grad.cpu(), o_leaves[k].grad)))
