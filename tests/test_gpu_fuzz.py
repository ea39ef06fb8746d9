"""Randomised configurations of the fused loss against the live oracle (GPU): number of scales, source
sets (mono / stereo / mixed), flags, batch sizes on both sides of the kernels' internal limits, image sizes
that are odd multiples of the pyramid divisor, jittered intrinsics.  Loss parity 1e-5 (north_star); gradients
bounded in relative L2 (tiny images: a single bilinear-cell / argmin flip weighs a lot) and
checked per sample so that no sample is left without its gradient."""
import os
import random

import pytest
import torch

from helpers import rel_l2
from oracle import view_synthesis as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

FRAME_SETS = [[0, -1], [0, 1], [0, -1, 1], [0, "s"], [0, -1, 1, "s"], [0, 1, "s"], [0, -2, -1, 1, 2], [0, -1, 1, 2, "s"]]


def _cases(n=14, seed=2024):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        S = rng.choice([1, 2, 3, 4, 4])
        div = 1 << (S - 1)
        H = div * rng.choice([3, 4, 5, 7]) * (2 if div < 8 else 1)
        W = div * rng.choice([5, 6, 9, 11]) * (2 if div < 8 else 1)
        H, W = max(H, 2 * div, 8), max(W, 2 * div, 8)
        out.append(dict(S=S, H=H, W=W, B=rng.choice([1, 2, 3, 17, 33]), fids=rng.choice(FRAME_SETS),
                        avg=rng.random() < 0.25, noauto=rng.random() < 0.25, no_ssim=rng.random() < 0.2,
                        kind=rng.choice(["iid", "structured"]), jitter=rng.random() < 0.5, seed=100 + i))
    return out


# MD2_FUZZ_N / MD2_FUZZ_SEED widen the sweep for a one-off hunt (default: 14 cases, fixed seed)
@pytest.mark.parametrize("c", _cases(int(os.environ.get("MD2_FUZZ_N", "14")), int(os.environ.get("MD2_FUZZ_SEED", "2024"))), ids=lambda c: "S%d_%dx%d_B%d_%s%s%s%s" % (
    c["S"], c["H"], c["W"], c["B"], "".join(str(f) for f in c["fids"]), "_avg" if c["avg"] else "",
    "_noauto" if c["noauto"] else "", "_l1" if c["no_ssim"] else ""))
def test_random_configuration_matches_oracle(c):
    from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
    from monodepth2_b200.synthetic import make_batch
    S, H, W, B, fids = c["S"], c["H"], c["W"], c["B"], c["fids"]
    if c["kind"] == "structured" and (H < 16 or W < 16):
        c = dict(c, kind="iid")
    n_src = len(fids) - 1
    n_id = 0 if c["noauto"] else (1 if c["avg"] else n_src)
    inputs, outputs, pose, noise = make_batch(B, H, W, fids, S, c["seed"], c["kind"], jitter_K=c["jitter"],
                                              n_id=max(n_id, 1))
    scales = tuple(range(S))
    cfg = O.OracleConfig(height=H, width=W, scales=scales, frame_ids=tuple(fids), avg_reprojection=c["avg"],
                         disable_automasking=c["noauto"], no_ssim=c["no_ssim"])
    o_outs = {k: v.clone().requires_grad_(True) for k, v in outputs.items()}
    o_losses = O.view_synthesis_loss(dict(inputs), o_outs, cfg, noise if n_id else None)
    o_losses["loss"].backward()
    # the reference's own fp32-vs-fp64 noise on this very case (protocol P3 of SURVEY.md 8c)
    d_outs = {k: v.double().clone().requires_grad_(True) for k, v in outputs.items()}
    d_losses = O.view_synthesis_loss({k: v.double() for k, v in inputs.items()}, d_outs, cfg,
                                     [n.double() for n in noise] if n_id else None)
    d_losses["loss"].backward()

    plan = LossPlan(B, H, W, fids, scales=list(scales), avg_reprojection=c["avg"], disable_automasking=c["noauto"],
                    no_ssim=c["no_ssim"])
    ins = {k: v.to(DEV) for k, v in inputs.items()}
    outs = {k: v.to(DEV).requires_grad_(True) for k, v in outputs.items()}
    losses = view_synthesis_loss(plan, ins, outs, [n.to(DEV) for n in noise] if n_id else None)
    losses["loss"].backward()
    torch.cuda.synchronize()

    for key in ["loss"] + ["loss/%d" % s for s in scales]:
        ref = float(o_losses[key].detach())
        assert abs(float(losses[key].detach()) - ref) <= 1e-5 * abs(ref), key
    # P3: no worse than 1.5 x the reference's own fp32 noise against fp64, plus the weight of a couple of single
    # discrete-decision flips on an image this small (a flip moves a handful of the B*H*W per-pixel terms by their
    # own magnitude: ~ sqrt(k / (B H W)) in relative L2; VERDICT r1 item 4c replaced the fixed 0.25-0.6 bounds)
    flip_floor = 2.0 / (B * H * W) ** 0.5
    for k in outputs:
        g, r = outs[k].grad.cpu(), o_outs[k].grad
        assert torch.isfinite(g).all(), k
        ref_noise = rel_l2(r, d_outs[k].grad)
        mine = rel_l2(g, d_outs[k].grad)
        assert mine <= 1.5 * ref_noise + flip_floor, (k, mine, ref_noise, flip_floor)
        num = (g - r).flatten(1).norm(dim=1)
        den = r.flatten(1).norm(dim=1)
        assert bool((num <= 0.9 * den + 1e-3 * den.max() + 1e-12).all()), (k, (num / (den + 1e-30)).tolist())


def test_bounds_checked_sweep_of_random_configurations():
    """VERDICT r1 item 4d (compute-sanitizer is closed on this pool): a 120-case sweep of random configurations
    through the debug build of the library, in which every global-memory index of the marching path is checked
    inside the kernels (-DMD2_BOUNDS_CHECK): no index may fall outside its tensor, and every gradient is finite."""
    from dbg_driver import debug_lib, oob_count
    from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
    from monodepth2_b200.synthetic import make_batch
    lib = debug_lib()
    oob_count(reset=True)
    for c in _cases(120, 77):
        S, H, W, B, fids = c["S"], c["H"], c["W"], min(c["B"], 5), c["fids"]
        kind = "iid" if (H < 16 or W < 16) else c["kind"]
        n_id = 0 if c["noauto"] else (1 if c["avg"] else len(fids) - 1)
        inputs, outputs, pose, noise = make_batch(B, H, W, fids, S, c["seed"], kind, jitter_K=c["jitter"], n_id=max(n_id, 1))
        plan = LossPlan(B, H, W, fids, scales=list(range(S)), avg_reprojection=c["avg"], disable_automasking=c["noauto"],
                        no_ssim=c["no_ssim"], rows_per_segment=random.Random(c["seed"]).choice([0, 0, 8, 16, 24]))
        plan.lib = lib
        ins = {k: v.to(DEV) for k, v in inputs.items()}
        outs = {k: v.to(DEV).requires_grad_(True) for k, v in outputs.items()}
        losses = view_synthesis_loss(plan, ins, outs, [n.to(DEV) for n in noise] if n_id else None)
        losses["loss"].backward()
        for k, v in outs.items():
            assert torch.isfinite(v.grad).all(), (c, k)
    torch.cuda.synchronize()
    assert oob_count() == 0
