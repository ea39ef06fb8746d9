import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from monodepth2_b200.synthetic import make_batch
from monodepth2_b200.fused_loss import LossPlan, GraphedLoss
dev = torch.device("cuda:0")
B, H, W, fids = 12, 192, 640, [0, -1, 1]
plan = LossPlan(B, H, W, fids)
q = lambda x: (x * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
inputs, outputs, pose, noise = make_batch(B, H, W, fids, 4, 0, "iid")
pin_in = {("color", f, 0): q(inputs[("color", f, 0)]).pin_memory() for f in fids}
for s in range(1, 4):
    pin_in[("color", 0, s)] = q(inputs[("color", 0, s)]).pin_memory()
for k in [("K", 0), ("inv_K", 0)]:
    pin_in[k] = inputs[k].pin_memory()
pin_out = {k: v.pin_memory() for k, v in outputs.items()}
g = GraphedLoss(plan, {k: v.to(dev) for k, v in pin_in.items()}, {k: v.to(dev) for k, v in pin_out.items()})
flat, s_in, s_out = g.staging()
for k, v in pin_in.items(): s_in[k].copy_(v)
for k, v in pin_out.items(): s_out[k].copy_(v)
print("bytes", flat.numel(), "pinned", flat.is_pinned())
def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
raw = torch.empty(flat.numel(), dtype=torch.uint8, device=dev)
flat2 = torch.empty(flat.numel(), dtype=torch.uint8).pin_memory(); flat2.copy_(flat)
import ctypes
for rep in range(3):
    print("rep", rep,
          "load_staged %.3f" % timeit(lambda: g.load_staged(flat), 100),
          "per-tensor %.3f" % timeit(lambda: g.load(pin_in, pin_out), 100),
          "raw(flat) %.3f" % timeit(lambda: raw.copy_(flat, non_blocking=True), 100),
          "raw(flat2) %.3f" % timeit(lambda: raw.copy_(flat2, non_blocking=True), 100),
          "_flat<-flat2 %.3f" % timeit(lambda: g._flat.copy_(flat2, non_blocking=True), 100),
          "replay %.3f" % timeit(lambda: g.run(), 100))
cs = torch.cuda.Stream()
def overlapped():
    with torch.cuda.stream(cs):
        raw.copy_(flat, non_blocking=True)
    g.run()
print("copy(raw) on side stream + replay %.3f" % timeit(overlapped, 100))
