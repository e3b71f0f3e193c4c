"""Dilated ResNet-101 backbone on the fused sm_100a kernels.

Parameter names, shapes, initialisation and the output-stride handling follow zs3/modeling/backbone/resnet.py
(so state_dict()s are interchangeable); the tree is generated from the tables in modeling/_build.py and every
block executes as ONE fused autograd node (functional.BottleneckFn) on NHWC bf16 activations.
"""
import os

import torch
import torch.nn as nn

from ... import functional as ZF
from .. import _build as B


class Bottleneck(nn.Module):
    """1x1 reduce -> 3x3 (stride / dilation) -> 1x1 expand (+ projection shortcut), reference resnet.py:9-53"""
    expansion = B.EXPANSION

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, BatchNorm=None):
        super().__init__()
        self.inplanes, self.planes = inplanes, planes
        widths = ((inplanes, planes, 1, 1, 1), (planes, planes, 3, stride, dilation),
                  (planes, planes * self.expansion, 1, 1, 1))
        for i, (cin, cout, k, s, d) in enumerate(widths, start=1):
            setattr(self, f"conv{i}", B.conv(cin, cout, k, s, d))
            setattr(self, f"bn{i}", BatchNorm(cout))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        """x: NHWC bf16 [N, H, W, cpad(inplanes)]"""
        return ZF.bottleneck(self, x)


_DP_CUT = os.environ.get("ZS3_DP_CUT", "1") == "1"   # same default as parallel.DataParallelTrainer

class ResNet(nn.Module):
    """stem + 4 residual stages; returns (stage-4 features, stage-1 'low level' features), reference resnet.py:56-226"""

    def __init__(self, block, layers, output_stride, BatchNorm, pretrained=True, imagenet_pretrained_path=""):
        super().__init__()
        strides, dilations = B.geometry(B.STAGE_GEOMETRY, output_stride)
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = BatchNorm(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        width = 64
        for idx, ((planes, _), depth) in enumerate(zip(B.RESNET101_STAGES, layers)):
            last = idx == len(layers) - 1
            # the last stage is the "multi-grid" unit: one block per MULTI_GRID entry, dilation scaled by it
            rates = [m * dilations[idx] for m in B.MULTI_GRID] if last else [dilations[idx]] * depth
            stage, width = self._stage(block, width, planes, strides[idx], rates, BatchNorm)
            setattr(self, f"layer{idx + 1}", stage)
        B.init_fan_out_(self)
        if pretrained:
            self._load_pretrained_model(imagenet_pretrained_path)

    @staticmethod
    def _stage(block, width, planes, stride, rates, BatchNorm):
        out_width = planes * block.expansion
        blocks = []
        for i, rate in enumerate(rates):
            s = stride if i == 0 else 1
            shortcut = None
            if i == 0 and (s != 1 or width != out_width):
                shortcut = nn.Sequential(B.conv(width, out_width, 1, s), BatchNorm(out_width))
            blocks.append(block(width, planes, s, rate, shortcut, BatchNorm))
            width = out_width
        return nn.Sequential(*blocks), width

    def forward(self, input):
        """input: NCHW fp32 CUDA image batch.  Returns (x, low_level_feat) as NHWC bf16 tensors."""
        if not input.is_cuda:
            raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        x = ZF.Stem.apply(self.conv1, self.bn1, self.maxpool, self.conv1.weight, self.bn1.weight, self.bn1.bias, input)
        low_level_feat = x = self.layer1(x)
        x = self.layer2(x)
        # the activations that separate {stem, layer1, layer2} from the rest of the network: the data-parallel
        # runtime cuts the backward pass here to overlap the gradient all-reduce of everything above the cut
        # (97 % of the parameters) with the backward of everything below it
        # (default; ZS3_DP_CUT=0 disables it, see parallel.py.  The decoder then consumes an ALIAS of layer1's output:
        # low_level_feat itself is upstream of x, and a cut must be an antichain -- with the alias, the decoder's
        # gradient arrives at a node of its own.)
        # Only a multi-rank DataParallelTrainer asks for it (it sets `expose_cut`): single-GPU graphs stay without the alias.
        self.last_cut = None
        if _DP_CUT and getattr(self, "expose_cut", False):
            low_level_feat = low_level_feat.view_as(low_level_feat)
            self.last_cut = (x, low_level_feat)
        x = self.layer4(self.layer3(x))
        return x, low_level_feat

    def _init_weight(self):
        B.init_fan_out_(self)

    def _load_pretrained_model(self, imagenet_pretrained_path):
        """ImageNet checkpoint keys carry a 'module.' prefix (7 characters, reference resnet.py:216-226)"""
        own = self.state_dict()
        loaded = torch.load(imagenet_pretrained_path)["state_dict"]
        own.update({k[7:]: v for k, v in loaded.items() if k[7:] in own})
        self.load_state_dict(own)


def ResNet101(output_stride, BatchNorm, pretrained=True, imagenet_pretrained_path=""):
    depths = [d for _, d in B.RESNET101_STAGES]
    return ResNet(Bottleneck, depths, output_stride, BatchNorm, pretrained=pretrained,
                  imagenet_pretrained_path=imagenet_pretrained_path)


BACKBONES = {"resnet101": ResNet101}   # the reference ships exactly one (zs3/modeling/backbone/__init__.py:4-12)


def build_backbone(output_stride, BatchNorm, pretrained=True, imagenet_pretrained_path="", name="resnet101"):
    return BACKBONES[name](output_stride, BatchNorm, pretrained=pretrained,
                           imagenet_pretrained_path=imagenet_pretrained_path)
