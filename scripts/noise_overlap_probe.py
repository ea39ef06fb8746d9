"""How much of the four torch.randn draws hides behind the fused call when they are drawn for the NEXT step on a side
stream inside the same CUDA graph (development probe for GraphedLoss(prefetch_noise=True))."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from monodepth2_b200.synthetic import make_batch
from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
B, H, W, fids = 12, 192, 640, [0, -1, 1]
dev = torch.device("cuda:0")
batches = []
for i in range(4):
    inputs, outputs, pose, noise = make_batch(B, H, W, fids, seed=i)
    batches.append(({k: v.to(dev) for k, v in inputs.items()}, {k: v.to(dev).requires_grad_(True) for k, v in outputs.items()}))
plan = LossPlan(B, H, W, fids)
shape = (B, plan.n_id, H, W)
N = [[torch.randn(shape, device=dev) for _ in range(4)] for _ in range(2)]
side = torch.cuda.Stream()

def step(mode, ins, outs):
    for v in outs.values():
        v.grad = None
    if mode == "public":
        l = view_synthesis_loss(plan, ins, outs)
    elif mode == "predrawn":
        l = view_synthesis_loss(plan, ins, outs, noise=N[0])
    else:
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for t in N[1]:
                t.normal_()
        l = view_synthesis_loss(plan, ins, outs, noise=N[0])
    l["loss"].backward()
    if mode == "prefetch":
        torch.cuda.current_stream().wait_stream(side)
    return l

for mode in ["public", "predrawn", "prefetch", "public", "predrawn", "prefetch"]:
    graphs = []
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for ins, outs in batches:
            step(mode, ins, outs)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    for ins, outs in batches:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step(mode, ins, outs)
        graphs.append(g)
    for _ in range(3):
        for g in graphs:
            g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 25
    e0.record()
    for _ in range(n):
        for g in graphs:
            g.replay()
    e1.record(); torch.cuda.synchronize()
    print(mode, "ms/step %.4f" % (e0.elapsed_time(e1) / (n * len(graphs))))
