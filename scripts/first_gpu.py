import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, numpy as np
from helpers import Golden, golden_cases, rel_l2
from gpu_driver import run_cuda
print(torch.cuda.get_device_name(0))
for name in golden_cases():
    g = Golden(name); z = g.z
    r = run_cuda(g, rows_per_segment=16)
    print(name, "loss", float(r["losses"]["loss"]), float(z["loss"]))
    for s in range(4):
        print("   s%d grad_disp relL2 %.3e" % (s, rel_l2(r["leaves"][("disp", s)].grad.cpu(), z["grad_disp__%d" % s])))
    for f in g.frame_ids[1:]:
        if f != 's': print("   gradT", f, rel_l2(r["leaves"][("T", f)].grad.cpu(), z["grad_cam_T_cam__%s" % f]))
# timing at full size
from monodepth2_b200.synthetic import make_batch
from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
for rows in (16, 24, 32, 48, 64):
    B,H,W = 12,192,640
    inputs, outputs, pose, noise = make_batch(B,H,W)
    dev='cuda:0'
    inputs={k:v.to(dev) for k,v in inputs.items()}
    outs={k:v.to(dev).requires_grad_(True) for k,v in outputs.items()}
    noise=[n.to(dev) for n in noise]
    plan=LossPlan(B,H,W,[0,-1,1],rows_per_segment=rows)
    for i in range(5):
        l=view_synthesis_loss(plan,inputs,outs,noise); l["loss"].backward()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    n=20
    e0.record()
    for i in range(n):
        l=view_synthesis_loss(plan,inputs,outs,noise)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n
    print("rows",rows,"fwd+bwd kernels only ms/step", ms, "frames/s", B/ms*1e3, "loss", float(l["loss"]))
