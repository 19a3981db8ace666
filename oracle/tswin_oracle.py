"""Restatement of the segmentation model AROUND the Swin head, for the full-model integration tests.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``TswinPlus`` below restates, in plain torch,

  * TswinPlus                      seg18/net/Ours/base18.py:52-108   (the call into the head: :94)
  * ResNet_BasicBlock_OS8 (18)     seg18/net/Ours/resnet.py:10-52, 92-136 (torchvision resnet18 trunk + dilated layer4/5)
  * ASPP                           seg18/net/Ours/ASPP.py:7-60

with the same attribute names, so that its ``state_dict`` has exactly the reference's keys and shapes
(``tests/golden/tswinplus.npz`` records the key list of the reference model; ``tests/test_oracle_golden.py``
checks it).  The Swin head is a constructor argument: the tests plug in ``stswincl_b200.swin`` (the product) --
that is the ``model.swin = ...`` swap of INTEGRATION.md -- or, on the CPU, the functional oracle.

Weights do not travel (107 M parameters): both sides fill the model from ``synth_state_dict``, a deterministic
function of the parameter NAMES and shapes only (independent of construction order and RNG consumption).
"""
from __future__ import annotations

import hashlib
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _make_layer(in_channels, channels, num_blocks, stride=1, dilation=1):      # resnet.py:10-20
    blocks = []
    for s in [stride] + [1] * (num_blocks - 1):
        blocks.append(BasicBlock(in_channels, channels, stride=s, dilation=dilation))
        in_channels = channels
    return nn.Sequential(*blocks)


class BasicBlock(nn.Module):                                                    # resnet.py:22-52
    def __init__(self, in_channels, channels, stride=1, dilation=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, channels, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, 3, stride=1, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(channels)
        if stride != 1 or in_channels != channels:
            self.downsample = nn.Sequential(nn.Conv2d(in_channels, channels, 1, stride=stride, bias=False), nn.BatchNorm2d(channels))
        else:
            self.downsample = nn.Sequential()

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + self.downsample(x))


class ResNet18_OS8(nn.Module):                                                  # resnet.py:92-136
    def __init__(self):
        super().__init__()
        import torchvision.models as models
        trunk = models.resnet18()
        self.resnet = nn.Sequential(*list(trunk.children())[:-4])             # conv1 .. layer2 (stride 8)
        self.layer4 = _make_layer(128, 256, 2, stride=1, dilation=2)
        self.layer5 = _make_layer(256, 512, 2, stride=1, dilation=4)

    def forward(self, x):
        return self.layer5(self.layer4(self.resnet(x)))


class ASPP(nn.Module):                                                          # ASPP.py:7-60
    def __init__(self, num_classes):
        super().__init__()
        c = 1024
        self.conv_1x1_1 = nn.Conv2d(c, c // 2, 1)
        self.bn_conv_1x1_1 = nn.BatchNorm2d(c // 2)
        self.conv_3x3_1 = nn.Conv2d(c, c // 2, 3, padding=6, dilation=6)
        self.bn_conv_3x3_1 = nn.BatchNorm2d(c // 2)
        self.conv_3x3_2 = nn.Conv2d(c, c // 2, 3, padding=12, dilation=12)
        self.bn_conv_3x3_2 = nn.BatchNorm2d(c // 2)
        self.conv_3x3_3 = nn.Conv2d(c, c // 2, 3, padding=18, dilation=18)
        self.bn_conv_3x3_3 = nn.BatchNorm2d(c // 2)
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv_1x1_2 = nn.Conv2d(c, c // 2, 1)
        self.bn_conv_1x1_2 = nn.BatchNorm2d(c // 2)
        self.conv_1x1_3 = nn.Conv2d(c // 2 * 5, c // 2, 1)
        self.bn_conv_1x1_3 = nn.BatchNorm2d(c // 2)
        self.conv_1x1_4 = nn.Conv2d(c // 2, num_classes, 1)

    def forward(self, x):
        h, w = x.shape[2:]
        a = F.relu(self.bn_conv_1x1_1(self.conv_1x1_1(x)))
        b = F.relu(self.bn_conv_3x3_1(self.conv_3x3_1(x)))
        c = F.relu(self.bn_conv_3x3_2(self.conv_3x3_2(x)))
        d = F.relu(self.bn_conv_3x3_3(self.conv_3x3_3(x)))
        e = F.relu(self.bn_conv_1x1_2(self.conv_1x1_2(self.avg_pool(x))))
        e = F.interpolate(e, size=(h, w), mode="bilinear", align_corners=False)   # F.upsample(mode="bilinear")
        out = F.relu(self.bn_conv_1x1_3(self.conv_1x1_3(torch.cat([a, b, c, d, e], 1))))
        return self.conv_1x1_4(out)


def _proj(cin):
    return nn.Sequential(nn.Conv2d(cin, 48, 1, bias=False), nn.BatchNorm2d(48), nn.ReLU(inplace=True))


class TswinPlus(nn.Module):                                                     # base18.py:52-108
    def __init__(self, num_classes, swin: nn.Module):
        super().__init__()
        self.swin = swin                                                       # SwinTransformerLayerv5() (:57)
        self.resnet = ResNet18_OS8()
        self.aspp = ASPP(num_classes=256)
        self.project1, self.project2, self.project3 = _proj(512), _proj(512), _proj(1024)
        self.classifier = nn.Sequential(nn.Conv2d(400, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                        nn.Conv2d(256, num_classes, 1))

    def forward(self, x):
        b, t, _, h, w = x.shape
        tem = torch.cat([self.resnet(x[:, i]).unsqueeze(1) for i in range(t)], dim=1)     # :84-89
        res_output = tem[:, -1]
        tem1, tem2 = self.swin(tem)                                                        # :94 -- the hot path
        t1, t2 = tem1[:, -1], tem2[:, -1]
        aspp_output = self.aspp(t2)
        p1, p2, p3 = self.project1(res_output), self.project2(t1), self.project3(t2)
        p3 = F.interpolate(p3, size=p1.shape[2:], mode="bilinear", align_corners=False)
        aspp_output = F.interpolate(aspp_output, size=p1.shape[2:], mode="bilinear", align_corners=False)
        out = self.classifier(torch.cat([p1, p2, p3, aspp_output], dim=1))
        return F.interpolate(out, (h, w), mode="bilinear")


def synth_state_dict(reference_state: dict, seed: int = 0) -> dict:
    """Deterministic weights from names and shapes: every tensor is drawn from a generator seeded by a hash of its
    key.  Convolution / linear weights ~ N(0, 2 / fan_in) (activations keep their scale through the ReLU stacks),
    norm scales 1 + 0.1 N, biases 0.05 N, relative-position tables 0.5 N (a visible bias), BatchNorm running statistics
    mean 0.05 N / variance 1 + 0.1 U; integer buffers (relative_position_index, num_batches_tracked) and the {0,-100}
    attn_mask buffers are kept."""
    out = {}
    for k, v in reference_state.items():
        if not v.is_floating_point() or k.endswith("attn_mask"):
            out[k] = v.clone()
            continue
        g = torch.Generator().manual_seed(seed * 1_000_003 + int(hashlib.sha256(k.encode()).hexdigest()[:12], 16) % (2 ** 31))
        n = torch.randn(v.shape, generator=g)
        if k.endswith("relative_position_bias_table"):
            t = 0.5 * n
        elif k.endswith("running_var"):
            t = 1.0 + 0.1 * torch.rand(v.shape, generator=g)
        elif k.endswith("running_mean"):
            t = 0.05 * n
        elif v.dim() >= 2:
            fan_in = v[0].numel()
            t = n * math.sqrt(2.0 / fan_in)
        elif k.endswith("weight"):
            t = 1.0 + 0.1 * n
        else:
            t = 0.05 * n
        out[k] = t.to(v.dtype)
    return out


def make_clip(seed: int, B: int = 1, T: int = 4, H: int = 512, W: int = 640) -> torch.Tensor:
    """Synthetic clip ~ U(0, 1) (the loader divides by 255, Endovis2018_new.py:88-89), smooth in space so that the
    segmentation is not pure noise."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand(B * T, 3, H // 32, W // 32, generator=g)
    x = F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False) + 0.05 * torch.randn(B * T, 3, H, W, generator=g)
    return x.clamp_(0, 1).view(B, T, 3, H, W)


def make_targets(seed: int, B: int, H: int, W: int, classes: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    coarse = torch.randint(0, classes, (B, 1, H // 64, W // 64), generator=g).float()
    return F.interpolate(coarse, size=(H, W), mode="nearest").long().squeeze(1)


class OracleSwin(nn.Module):
    """The functional CPU oracle of the head (``swin_oracle.swin_layer_v5``) behind the reference's module
    interface: parameters / buffers are registered under the reference's state_dict keys (nested empty modules), the
    forward is ``forward(x [B,4,C,H,W]) -> (x3, x6)``.  Used to pin the restated caller model on the CPU."""

    def __init__(self, shapes: dict, dim=512, input_resolution=(64, 80), num_heads=4):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, tuple(input_resolution), num_heads
        for key, (shape, is_param, dtype) in shapes.items():
            mod = self
            parts = key.split(".")
            for p in parts[:-1]:
                if not hasattr(mod, p):
                    mod.add_module(p, nn.Module())
                mod = getattr(mod, p)
            t = torch.zeros(shape, dtype=dtype)
            if is_param:
                mod.register_parameter(parts[-1], nn.Parameter(t))
            else:
                mod.register_buffer(parts[-1], t)

    def forward(self, x):
        from . import swin_oracle as so
        params = dict(self.named_parameters())
        params.update(dict(self.named_buffers()))
        return so.swin_layer_v5(x, params, self.dim, self.input_resolution, self.num_heads)


def swin_shapes_from_keys(keys: dict) -> dict:
    """{'swin.<key>': shape} entries of the recorded reference key list -> OracleSwin's shape table."""
    out = {}
    for k, shape in keys.items():
        if not k.startswith("swin."):
            continue
        name = k[5:]
        is_buf = name.endswith("relative_position_index") or name.endswith("attn_mask")
        out[name] = (tuple(shape), not is_buf, torch.int64 if name.endswith("relative_position_index") else torch.float32)
    return out
