"""Quick device timing of the fused loss at full size (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monodepth2_b200.synthetic import make_batch
from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B, H, W = 12, 192, 640
inputs, outputs, pose, noise = make_batch(B, H, W)
dev = 'cuda:0'
inputs = {k: v.to(dev) for k, v in inputs.items()}
outs = {k: v.to(dev).requires_grad_(True) for k, v in outputs.items()}
noise = [x.to(dev) for x in noise]
plan = LossPlan(B, H, W, [0, -1, 1], rows_per_segment=rows)
for i in range(3):
    l = view_synthesis_loss(plan, inputs, outs, noise)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    l = view_synthesis_loss(plan, inputs, outs, noise)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("rows", rows, "ms/step", ms, "frames/s", B / ms * 1e3, "loss", float(l["loss"]))
