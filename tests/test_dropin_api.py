"""The drop-in boundary against the reference's own definitions (CPU; needs /root/reference, which exists in
the build container only - skipped elsewhere): same public names in layers.py, same constructor / call
signatures, the two Trainer methods the mixin overrides, and LossPlan.from_opt on the real option parser."""
import inspect
import os
import sys
import types

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not present on this box")


@pytest.fixture(scope="module")
def ref():
    for name, attrs in (("tensorboardX", {"SummaryWriter": object}), ("IPython", {"embed": lambda *a, **k: None}),
                        ("skimage", {}), ("skimage.transform", {})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        import layers as ref_layers
        import trainer as ref_trainer
        from options import MonodepthOptions
    finally:
        sys.path.remove(REF)
    return types.SimpleNamespace(layers=ref_layers, trainer=ref_trainer, Options=MonodepthOptions)


def _params(fn):
    return [p.name for p in inspect.signature(fn).parameters.values()]


def test_layers_exports_every_public_name_of_the_reference(ref):
    from monodepth2_b200 import layers as L
    ref_public = [n for n, v in vars(ref.layers).items()
                  if not n.startswith("_") and getattr(v, "__module__", None) == ref.layers.__name__]
    assert ref_public, "nothing found in the reference's layers.py"
    missing = [n for n in ref_public if not hasattr(L, n)]
    assert missing == []
    assert set(ref_public) <= set(L.__all__) | {"grid_sample_border"}


@pytest.mark.parametrize("name", ["BackprojectDepth", "Project3D", "SSIM", "ConvBlock", "Conv3x3"])
def test_module_signatures_match(ref, name):
    from monodepth2_b200 import layers as L
    a, b = getattr(ref.layers, name), getattr(L, name)
    assert _params(a.__init__) == _params(b.__init__), name
    assert _params(a.forward) == _params(b.forward), name


@pytest.mark.parametrize("name", ["disp_to_depth", "transformation_from_parameters", "get_translation_matrix",
                                  "rot_from_axisangle", "get_smooth_loss", "upsample", "compute_depth_errors"])
def test_function_signatures_match(ref, name):
    from monodepth2_b200 import layers as L
    assert _params(getattr(ref.layers, name)) == _params(getattr(L, name)), name


def test_mixin_overrides_exactly_the_two_hot_path_methods(ref):
    from monodepth2_b200.fused_loss import FusedLossMixin
    T = ref.trainer.Trainer
    for m in ("generate_images_pred", "compute_losses"):
        assert _params(getattr(T, m)) == _params(getattr(FusedLossMixin, m)), m

    class FusedTrainer(FusedLossMixin, T):
        pass
    assert FusedTrainer.generate_images_pred is FusedLossMixin.generate_images_pred
    assert FusedTrainer.compute_losses is FusedLossMixin.compute_losses
    assert FusedTrainer.process_batch is T.process_batch            # the caller stays the reference's
    assert FusedTrainer.predict_poses is T.predict_poses
    overridden = [n for n, v in vars(FusedLossMixin).items() if callable(v) and not n.startswith("_")]
    assert sorted(overridden) == ["compute_losses", "generate_images_pred"]


@pytest.mark.parametrize("flags", [[], ["--use_stereo"], ["--avg_reprojection"], ["--disable_automasking"],
                                   ["--no_ssim"], ["--v1_multiscale"], ["--pose_model_type", "posecnn"],
                                   ["--disable_automasking", "--predictive_mask"],
                                   ["--height", "320", "--width", "1024"], ["--frame_ids", "0", "--use_stereo"]])
def test_plan_from_the_reference_option_parser(ref, flags, monkeypatch):
    from monodepth2_b200 import _capi
    from monodepth2_b200.fused_loss import LossPlan
    if not os.path.exists(_capi.LIB_PATH):
        pytest.skip("libmd2loss.so not built")
    monkeypatch.setattr(sys, "argv", ["train.py"] + flags)
    opt = ref.Options().parse()
    if opt.use_stereo:
        opt.frame_ids.append("s")                                    # trainer.py:51-52
    plan = LossPlan.from_opt(opt)
    assert (plan.batch_size, plan.height, plan.width) == (opt.batch_size, opt.height, opt.width)
    assert plan.frame_ids == opt.frame_ids and plan.scales == opt.scales
    assert plan.automask == (not opt.disable_automasking)
    assert plan.avg_reprojection == opt.avg_reprojection and plan.no_ssim == opt.no_ssim
    assert plan.v1_multiscale == opt.v1_multiscale and plan.predictive_mask == opt.predictive_mask
    assert plan.posecnn == (opt.pose_model_type == "posecnn")
    assert (plan.min_depth, plan.max_depth, plan.disparity_smoothness) == (opt.min_depth, opt.max_depth,
                                                                          opt.disparity_smoothness)


def test_predictive_mask_without_disable_automasking_raises_like_the_reference(ref, monkeypatch):
    from monodepth2_b200 import _capi
    from monodepth2_b200.fused_loss import LossPlan
    if not os.path.exists(_capi.LIB_PATH):
        pytest.skip("libmd2loss.so not built")
    monkeypatch.setattr(sys, "argv", ["train.py", "--predictive_mask"])
    opt = ref.Options().parse()
    with pytest.raises(RuntimeError):                                # trainer.py:90-92 asserts the same
        LossPlan.from_opt(opt)
