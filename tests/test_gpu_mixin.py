"""FusedLossMixin the way INTEGRATION.md section 2 uses it: an object with the reference's `opt` namespace calls
generate_images_pred(inputs, outputs) then compute_losses(inputs, outputs) (trainer.py:257-258) and
losses["loss"].backward() (trainer.py:208); the tie-break noise comes from the CUDA generator as in
trainer.py:468-469, so seeding it reproduces the oracle's draws.  Also the logging-step side outputs
(`md2_side`) and the no_grad validation mode (trainer.py:330-331)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import rel_l2
from oracle import view_synthesis as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _opt(B, H, W, frame_ids, **kw):
    d = dict(batch_size=B, height=H, width=W, frame_ids=list(frame_ids), scales=[0, 1, 2, 3], min_depth=0.1,
             max_depth=100.0, disparity_smoothness=1e-3, avg_reprojection=False, disable_automasking=False,
             no_ssim=False, v1_multiscale=False, predictive_mask=False, pose_model_type="separate_resnet")
    d.update(kw)
    return SimpleNamespace(**d)


def _trainer(opt):
    from monodepth2_b200.fused_loss import FusedLossMixin

    class T(FusedLossMixin):
        def __init__(self, opt):
            self.opt = opt
    return T(opt)


@pytest.mark.parametrize("fids,kw", [([0, -1, 1], {}), ([0, -1, 1, "s"], {}), ([0, -1, 1], {"avg_reprojection": True})])
def test_mixin_matches_oracle_with_generator_noise(fids, kw):
    from monodepth2_b200.synthetic import make_batch
    B, H, W = 3, 64, 96
    inputs, outputs, pose, _ = make_batch(B, H, W, fids, 4, 51, "structured")
    n_src = len(fids) - 1
    n_id = 1 if kw.get("avg_reprojection") else n_src
    t = _trainer(_opt(B, H, W, fids, **kw))
    t.md2_side = {"depth_scales": [0], "color_scales": [0], "mask_scales": [0, 1, 2, 3]}
    ins = {k: v.to(DEV) for k, v in inputs.items()}
    outs = {k: v.to(DEV).requires_grad_(True) for k, v in outputs.items()}
    leaves = dict(outs)
    torch.manual_seed(123)
    t.generate_images_pred(ins, outs)
    losses = t.compute_losses(ins, outs)
    losses["loss"].backward()
    assert sorted(losses) == ["loss", "loss/0", "loss/1", "loss/2", "loss/3"]
    # the same four draws, in the same order, for the oracle
    torch.manual_seed(123)
    noise = [torch.randn((B, n_id, H, W), device=DEV).cpu() for _ in range(4)]
    cfg = O.OracleConfig(height=H, width=W, frame_ids=tuple(fids), **kw)
    o_outs = {k: v.clone().requires_grad_(True) for k, v in outputs.items()}
    o_leaves = dict(o_outs)
    o_losses = O.view_synthesis_loss(dict(inputs), o_outs, cfg, noise)
    o_losses["loss"].backward()
    for k in losses:
        ref = float(o_losses[k].detach())
        assert abs(float(losses[k].detach()) - ref) <= 1e-5 * abs(ref), k
    for k in leaves:
        assert rel_l2(leaves[k].grad.cpu(), o_leaves[k].grad) < 0.1, k
    # what Trainer.log / compute_depth_losses read on logging steps (trainer.py:504,553-572)
    np.testing.assert_allclose(outs[("depth", 0, 0)].cpu().numpy(), o_outs[("depth", 0, 0)].detach().numpy(), rtol=2e-6)
    for f in fids[1:]:
        np.testing.assert_allclose(outs[("color", f, 0)].cpu().numpy(), o_outs[("color", f, 0)].detach().numpy(), atol=5e-5)
    for s in range(4):
        m = outs["identity_selection/%d" % s].cpu()
        assert m.shape == (B, H, W)
        assert float((m != o_outs["identity_selection/%d" % s]).float().mean()) <= 1e-3


def test_mixin_under_no_grad_is_forward_only():
    from monodepth2_b200.synthetic import make_batch
    B, H, W, fids = 2, 48, 80, [0, -1, 1]
    inputs, outputs, pose, _ = make_batch(B, H, W, fids, 4, 52, "structured")
    t = _trainer(_opt(B, H, W, fids))
    ins = {k: v.to(DEV) for k, v in inputs.items()}
    outs = {k: v.to(DEV).requires_grad_(True) for k, v in outputs.items()}
    with torch.no_grad():                      # Trainer.val, trainer.py:330-331
        torch.manual_seed(5)
        t.generate_images_pred(ins, outs)
        l0 = t.compute_losses(ins, outs)
    assert not l0["loss"].requires_grad
    torch.manual_seed(5)
    t.generate_images_pred(ins, outs)
    l1 = t.compute_losses(ins, outs)
    assert l1["loss"].requires_grad
    assert abs(float(l0["loss"]) - float(l1["loss"].detach())) <= 2e-6 * abs(float(l0["loss"]))


def test_logging_step_materialises_side_outputs_on_demand():
    """Trainer.run_epoch decides to log after process_batch (trainer.py:213-227), so log() and
    compute_depth_losses() must find outputs[("depth",0,0)], ("color",f,0) and "identity_selection/s" although the
    training call produced none.  The base class below reads exactly the keys trainer.py:504,553-572 read."""
    from monodepth2_b200.fused_loss import FusedLossMixin
    from monodepth2_b200.synthetic import make_batch
    B, H, W, fids = 2, 48, 80, [0, -1, 1]
    seen = {}

    class RefLike:
        def compute_depth_losses(self, inputs, outputs, losses):
            seen["depth"] = outputs[("depth", 0, 0)]                       # trainer.py:504

        def log(self, mode, inputs, outputs, losses):
            for f in self.opt.frame_ids[1:]:
                seen[("color", f)] = outputs[("color", f, 0)]              # trainer.py:561-563
            for s in self.opt.scales:
                seen[("mask", s)] = outputs["identity_selection/{}".format(s)]   # trainer.py:570-572

    class T(FusedLossMixin, RefLike):
        md2_fused_metrics = False          # this test is about the side outputs the base-class methods read

        def __init__(self, opt):
            self.opt = opt

    def step(eager):
        inputs, outputs, pose, _ = make_batch(B, H, W, fids, 4, 53, "structured")
        t = T(_opt(B, H, W, fids))
        if eager:
            t.md2_side = {"depth_scales": [0], "color_scales": [0], "mask_scales": [0, 1, 2, 3]}
        ins = {k: v.to(DEV) for k, v in inputs.items()}
        outs = {k: v.to(DEV).requires_grad_(True) for k, v in outputs.items()}
        torch.manual_seed(7)
        t.generate_images_pred(ins, outs)
        losses = t.compute_losses(ins, outs)
        losses["loss"].backward()
        return t, ins, outs, losses

    t, ins, outs, losses = step(eager=False)
    assert ("depth", 0, 0) not in outs and "identity_selection/0" not in outs
    seen.clear()
    t.compute_depth_losses(ins, outs, losses)
    t.log("train", ins, outs, losses)
    lazy = {k: v.clone() for k, v in seen.items()}
    t2, ins2, outs2, losses2 = step(eager=True)
    seen.clear()
    t2.compute_depth_losses(ins2, outs2, losses2)
    t2.log("train", ins2, outs2, losses2)
    assert abs(float(losses["loss"].detach()) - float(losses2["loss"].detach())) <= 1e-7 * abs(float(losses2["loss"].detach()))
    assert sorted(map(str, lazy)) == sorted(map(str, seen))
    for k in seen:
        assert torch.equal(lazy[k], seen[k]), k                            # same noise -> same masks, same images
    # and the validation path (trainer.py:320-339: no_grad, then compute_depth_losses + log)
    with torch.no_grad():
        inputs, outputs, pose, _ = make_batch(B, H, W, fids, 4, 54, "structured")
        tv = T(_opt(B, H, W, fids))
        insv = {k: v.to(DEV) for k, v in inputs.items()}
        outsv = {k: v.to(DEV) for k, v in outputs.items()}
        tv.generate_images_pred(insv, outsv)
        lv = tv.compute_losses(insv, outsv)
        seen.clear()
        tv.compute_depth_losses(insv, outsv, lv)
        tv.log("val", insv, outsv, lv)
        assert seen["depth"].shape == (B, 1, H, W) and seen[("mask", 3)].shape == (B, H, W)


def test_graphed_loss_replays_match_the_eager_call():
    """GraphedLoss (public API): the fused call + backward captured in a CUDA graph over static buffers.  With
    --disable_automasking there is no tie-break noise, so a replay on freshly loaded inputs must reproduce the eager
    call on the same inputs exactly (per-pixel gradients; the scalar loss up to the order of its fp64 atomics)."""
    from monodepth2_b200.fused_loss import GraphedLoss, LossPlan, view_synthesis_loss
    from monodepth2_b200.synthetic import make_batch
    B, H, W, fids = 2, 64, 96, [0, -1, 1]
    plan = LossPlan(B, H, W, fids, disable_automasking=True)
    b0 = make_batch(B, H, W, fids, 4, 71, "structured")
    b1 = make_batch(B, H, W, fids, 4, 72, "iid")
    need = lambda ins: {k: v for k, v in ins.items() if k in [("color", f, 0) for f in fids] + [("color", 0, s) for s in range(1, 4)] + [("K", 0), ("inv_K", 0)]}
    g = GraphedLoss(plan, {k: v.to(DEV) for k, v in need(b0[0]).items()}, {k: v.to(DEV) for k, v in b0[1].items()})
    flat, s_in, s_out = g.staging()          # pinned staging buffer laid out like the static device buffers
    for n, batch in enumerate((b1, b0, b1, b0)):
        if n < 2:     # per-tensor copies
            g.load({k: v.pin_memory() for k, v in need(batch[0]).items()}, {k: v.pin_memory() for k, v in batch[1].items()})
        else:         # one copy of the whole step
            for k, v in need(batch[0]).items():
                s_in[k].copy_(v)
            for k, v in batch[1].items():
                s_out[k].copy_(v)
            g.load_staged(flat)
        if n % 2:
            loss = float(g.run().item())
        else:         # the non-blocking form: D2H copy of the loss put in flight with the replay, read later
            g.run_async()
            loss = g.read()
        ins = {k: v.to(DEV) for k, v in batch[0].items()}
        outs = {k: v.to(DEV).requires_grad_(True) for k, v in batch[1].items()}
        ref = view_synthesis_loss(plan, ins, outs)
        ref["loss"].backward()
        assert abs(loss - float(ref["loss"].detach())) <= 1e-7 * abs(loss)
        for s in range(4):
            assert torch.equal(g.grads[("disp", s)], outs[("disp", s)].grad), s


def test_mixin_compute_depth_losses_runs_the_fused_metrics():
    """trainer.py:498-526 through md2_depth_metrics: same keys, numpy values, equal to the oracle restatement."""
    from monodepth2_b200.fused_loss import FusedLossMixin
    from monodepth2_b200.synthetic import make_batch
    B, H, W, fids = 2, 48, 80, [0, -1, 1]
    inputs, outputs, pose, _ = make_batch(B, H, W, fids, 4, 55, "structured")

    class T(FusedLossMixin):
        depth_metric_names = ["de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"]

        def __init__(self, opt):
            self.opt = opt
    t = T(_opt(B, H, W, fids))
    ins = {k: v.to(DEV) for k, v in inputs.items()}
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(B, 1, 375, 1242, generator=g) * 60
    gt[torch.rand(gt.shape, generator=g) < 0.8] = 0
    ins["depth_gt"] = gt.to(DEV)
    outs = {k: v.to(DEV).requires_grad_(True) for k, v in outputs.items()}
    t.generate_images_pred(ins, outs)
    losses = t.compute_losses(ins, outs)
    t.compute_depth_losses(ins, outs, losses)
    ref = O.depth_metrics(outs[("depth", 0, 0)].detach().cpu(), gt)
    for i, k in enumerate(T.depth_metric_names):
        assert isinstance(losses[k], np.ndarray)
        assert abs(float(losses[k]) - float(ref[i])) <= 2e-5 * abs(float(ref[i])) + 1e-7, k


def test_calls_from_two_host_threads_on_two_streams():
    """The library's side streams and fork / join events are shared per device; the enqueue of a call runs under a
    lock (md2_kernels.cu: g_enqueue_mu), so two host threads, each with its own stream, plan and workspace, may call
    concurrently (ctypes drops the GIL during the call).  Same results as the serial run: per-pixel gradients bit for
    bit, losses up to the order of the fp64 atomic sums."""
    import threading
    from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
    from monodepth2_b200.synthetic import make_batch
    dev = "cuda:0"
    fids = [0, -1, 1]
    jobs = []
    for seed in (41, 42):
        inputs, outputs, _pose, noise = make_batch(4, 96, 320, fids, 4, seed, "structured", n_id=2)
        jobs.append(dict(plan=LossPlan(4, 96, 320, fids), ins={k: v.to(dev) for k, v in inputs.items()},
                         outs={k: v.to(dev).requires_grad_(True) for k, v in outputs.items()},
                         noise=[n.to(dev) for n in noise], stream=torch.cuda.Stream(device=dev)))

    def run(j, reps):
        with torch.cuda.stream(j["stream"]):
            for _ in range(reps):
                for v in j["outs"].values():
                    v.grad = None
                ls = view_synthesis_loss(j["plan"], j["ins"], j["outs"], j["noise"])
                ls["loss"].backward()
            j["stream"].synchronize()
        return float(ls["loss"]), {k: v.grad.clone() for k, v in j["outs"].items()}

    torch.cuda.synchronize()
    serial = [run(j, 1) for j in jobs]
    got = [None, None]

    def worker(i):
        got[i] = run(jobs[i], 25)
    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    for (l0, g0), (l1, g1) in zip(serial, got):
        assert abs(l0 - l1) <= 2e-6 * abs(l0)
        for k in g0:
            if k[0] == "disp":
                assert torch.equal(g0[k], g1[k]), k
            else:
                assert float((g0[k] - g1[k]).abs().max()) <= 1e-5 * float(g0[k].abs().max()), k
