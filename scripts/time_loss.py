"""Quick device timing of the fused loss at full size (development aid).
usage: time_loss.py [rows] [steps] [workload: mono|stereo|hires]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monodepth2_b200.synthetic import make_batch
from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
wl = sys.argv[3] if len(sys.argv) > 3 else "mono"
kind = sys.argv[4] if len(sys.argv) > 4 else "iid"
nograd = len(sys.argv) > 5 and sys.argv[5] == "nograd"
B, H, W = (12, 320, 1024) if wl == "hires" else (12, 192, 640)
fids = [0, -1, 1, "s"] if wl == "stereo" else ([0, -1] if wl == "one" else [0, -1, 1])
inputs, outputs, pose, noise = make_batch(B, H, W, fids, kind=kind, seed=5 if kind == 'structured' else 0)
dev = 'cuda:0'
inputs = {k: v.to(dev) for k, v in inputs.items()}
outs = {k: v.to(dev).requires_grad_(not nograd) for k, v in outputs.items()}
noise = [x.to(dev) for x in noise]
plan = LossPlan(B, H, W, fids, rows_per_segment=rows)
for i in range(3):
    l = view_synthesis_loss(plan, inputs, outs, noise)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    l = view_synthesis_loss(plan, inputs, outs, noise)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
import ctypes
plan.lib.md2_profile_enable(1)
mm = []
for i in range(10):
    l = view_synthesis_loss(plan, inputs, outs, noise)
    v = ctypes.c_float(0); plan.lib.md2_profile_march_ms(ctypes.byref(v)); mm.append(v.value)
plan.lib.md2_profile_enable(0)
print(os.path.basename(os.environ.get("MD2_LIB_PATH","libmd2loss")), os.environ.get("MD2_PAD_SMEM","-"), "nograd" if nograd else "grad", wl, kind, "rows", rows, "ms/step %.4f" % ms, "march_ms %.4f" % (sum(mm) / len(mm)), "frames/s %.0f" % (B / ms * 1e3), "loss", float(l["loss"].detach()))
