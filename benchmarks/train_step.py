"""Full training step around the view-synthesis loss (BASELINE.json configs 3-5).

    nets -> predict_poses -> [generate_images_pred + compute_losses] -> backward -> Adam   (+ DDP all-reduce)

What is measured is the *caller* of the hot path, `Trainer.process_batch` + `run_epoch`'s optimiser
step (/root/reference/trainer.py:193-260), on synthetic batches that are already resident in HBM
(the KITTI dataloader is out of scope).  The networks stay stock PyTorch/cuDNN modules, as
BASELINE.json's north_star says; /root/reference does not exist on the GPU box, so this file holds
benchmark STAND-INS with the reference's topology and parameter counts (ResNet-18/50 encoders from
torchvision, a 5-stage up-convolution depth decoder with skips and 4 sigmoid disparity heads, a
6-channel pose encoder and a 3-conv pose decoder, random init = `--weights_init scratch`).  They
are not part of the product package: nothing under monodepth2_b200/ imports this module.

The loss arm is selected by the caller:
  * "fused":     monodepth2_b200.fused_loss.view_synthesis_loss  (one C-ABI call, sm_100a kernels)
  * "reference": the oracle port on torch's CUDA kernels (bench.py --impl reference --mode train)
Data parallelism: one process per GPU, batch 12 per rank, torch DDP over NCCL for the network
gradients only; the loss kernels are per-sample and need no collective.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv3x3_reflect(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(cin, cout, 3))


class _UpBlock(nn.Module):
    """reflect-pad 3x3 conv + ELU (the decoder building block, layers.py:106-136)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = _conv3x3_reflect(cin, cout)

    def forward(self, x):
        return F.elu(self.conv(x), inplace=True)


class Encoder(nn.Module):
    """torchvision ResNet trunk returning the 5 feature maps (networks/resnet_encoder.py:62-98)."""

    def __init__(self, num_layers: int = 18, num_input_images: int = 1):
        super().__init__()
        import torchvision.models as tvm
        net = {18: tvm.resnet18, 50: tvm.resnet50}[num_layers](weights=None)
        if num_input_images > 1:
            net.conv1 = nn.Conv2d(3 * num_input_images, 64, 7, 2, 3, bias=False)
            nn.init.kaiming_normal_(net.conv1.weight, mode="fan_out", nonlinearity="relu")
        net.fc = nn.Identity()
        self.net = net
        self.channels = [64, 64, 128, 256, 512] if num_layers <= 34 else [64, 256, 512, 1024, 2048]

    def forward(self, x):
        n = self.net
        x = (x - 0.45) / 0.225
        f0 = n.relu(n.bn1(n.conv1(x)))
        f1 = n.layer1(n.maxpool(f0))
        f2 = n.layer2(f1)
        f3 = n.layer3(f2)
        f4 = n.layer4(f3)
        return [f0, f1, f2, f3, f4]


class DepthDecoder(nn.Module):
    """5 up-convolution stages with skip connections, sigmoid disparity heads at 4 scales
    (networks/depth_decoder.py:17-65)."""

    def __init__(self, enc_channels: Sequence[int], scales: Sequence[int] = (0, 1, 2, 3)):
        super().__init__()
        dec = [16, 32, 64, 128, 256]
        self.scales = list(scales)
        self.up0 = nn.ModuleList()
        self.up1 = nn.ModuleList()
        for i in range(5):
            cin = enc_channels[-1] if i == 4 else dec[i + 1]
            self.up0.append(_UpBlock(cin, dec[i]))
            self.up1.append(_UpBlock(dec[i] + (enc_channels[i - 1] if i > 0 else 0), dec[i]))
        self.heads = nn.ModuleDict({str(s): _conv3x3_reflect(dec[s], 1) for s in self.scales})

    def forward(self, feats):
        out = {}
        x = feats[-1]
        for i in range(4, -1, -1):
            x = F.interpolate(self.up0[i](x), scale_factor=2, mode="nearest")
            if i > 0:
                x = torch.cat([x, feats[i - 1]], 1)
            x = self.up1[i](x)
            if i in self.scales:
                out[("disp", i)] = torch.sigmoid(self.heads[str(i)](x))
        return out


class PoseDecoder(nn.Module):
    """1x1 squeeze, two 3x3 convs, 1x1 head, global mean, x0.01 (networks/pose_decoder.py:14-54);
    predicts 2 frames, only [:, 0] is used by the caller (trainer.py:289-295)."""

    def __init__(self, enc_channels: Sequence[int]):
        super().__init__()
        self.squeeze = nn.Conv2d(enc_channels[-1], 256, 1)
        self.c0 = nn.Conv2d(256, 256, 3, 1, 1)
        self.c1 = nn.Conv2d(256, 256, 3, 1, 1)
        self.c2 = nn.Conv2d(256, 12, 1)

    def forward(self, feats):
        x = F.relu(self.squeeze(feats[-1]))
        x = F.relu(self.c0(x))
        x = F.relu(self.c1(x))
        x = self.c2(x).mean(3).mean(2)
        x = 0.01 * x.view(-1, 2, 1, 6)
        return x[..., :3], x[..., 3:]


class Nets(nn.Module):
    """The four networks of `Trainer.models` (trainer.py:38-88) in one container so that a single
    DDP wrapper covers them; forward() = the network part of process_batch + predict_poses
    (trainer.py:248-255, 262-295) and returns disparities and pose parameters."""

    def __init__(self, frame_ids: Sequence, num_layers: int = 18):
        super().__init__()
        self.frame_ids = list(frame_ids)
        self.encoder = Encoder(num_layers, 1)
        self.depth = DepthDecoder(self.encoder.channels)
        self.pose_encoder = Encoder(num_layers, 2)
        self.pose = PoseDecoder(self.pose_encoder.channels)

    def forward(self, colors: Dict):
        out = self.depth(self.encoder(colors[0]))
        for f in self.frame_ids[1:]:
            if f == "s":
                continue
            pair = [colors[f], colors[0]] if f < 0 else [colors[0], colors[f]]     # trainer.py:276-279
            aa, tr = self.pose(self.pose_encoder(torch.cat(pair, 1)))
            out[("axisangle", 0, f)] = aa
            out[("translation", 0, f)] = tr
        return out


class TrainStep:
    """One optimisation step on a resident synthetic batch; `loss_fn(inputs, outputs) -> dict`."""

    def __init__(self, nets: nn.Module, frame_ids: Sequence, loss_fn, pose_fn, lr: float = 1e-4):
        self.nets, self.frame_ids, self.loss_fn, self.pose_fn = nets, list(frame_ids), loss_fn, pose_fn
        self.opt = torch.optim.Adam(nets.parameters(), lr)

    def __call__(self, inputs: Dict) -> torch.Tensor:
        colors = {f: inputs[("color", f, 0)] for f in self.frame_ids}       # color_aug = color copy
        outputs = self.nets(colors)
        for f in self.frame_ids[1:]:
            if f == "s" or self.pose_fn is None:     # pose_fn None: the loss builds T from the pose leaves itself
                continue
            outputs[("cam_T_cam", 0, f)] = self.pose_fn(outputs[("axisangle", 0, f)][:, 0],
                                                        outputs[("translation", 0, f)][:, 0], f < 0)
        losses = self.loss_fn(inputs, outputs)
        self.opt.zero_grad(set_to_none=True)
        losses["loss"].backward()
        self.opt.step()
        return losses["loss"].detach()


def parameter_count(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
