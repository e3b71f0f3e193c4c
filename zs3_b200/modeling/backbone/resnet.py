"""Dilated ResNet-101 backbone on the fused sm_100a kernels.

Module tree, parameter names, init and config handling follow zs3/modeling/backbone/resnet.py so that
state_dict()s are interchangeable; the forward runs on NHWC bf16 activations through
zs3_b200.functional (conv+BN+ReLU(+residual) fused Functions) instead of nn.Conv2d/BatchNorm calls.
"""
import math

import torch
import torch.nn as nn

from ... import functional as ZF
from ..sync_batchnorm.batchnorm import SynchronizedBatchNorm2d


class Bottleneck(nn.Module):
    """zs3/modeling/backbone/resnet.py:9-53"""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, BatchNorm=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = BatchNorm(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, dilation=dilation, padding=dilation,
                               bias=False)
        self.bn2 = BatchNorm(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = BatchNorm(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.inplanes, self.planes = inplanes, planes

    def forward(self, x):
        """x: NHWC bf16 [N,H,W,cpad(inplanes)]"""
        return ZF.bottleneck(self, x)  # one fused autograd node per block (see functional.BottleneckFn)


class ResNet(nn.Module):
    """zs3/modeling/backbone/resnet.py:56-226"""

    def __init__(self, block, layers, output_stride, BatchNorm, pretrained=True, imagenet_pretrained_path=""):
        self.inplanes = 64
        super().__init__()
        blocks = [1, 2, 4]
        if output_stride == 16:
            strides = [1, 2, 2, 1]
            dilations = [1, 1, 1, 2]
        elif output_stride == 8:
            strides = [1, 2, 1, 1]
            dilations = [1, 1, 2, 4]
        else:
            raise NotImplementedError

        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = BatchNorm(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

        self.layer1 = self._make_layer(block, 64, layers[0], stride=strides[0], dilation=dilations[0],
                                       BatchNorm=BatchNorm)
        self.layer2 = self._make_layer(block, 128, layers[1], stride=strides[1], dilation=dilations[1],
                                       BatchNorm=BatchNorm)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=strides[2], dilation=dilations[2],
                                       BatchNorm=BatchNorm)
        self.layer4 = self._make_MG_unit(block, 512, blocks=blocks, stride=strides[3], dilation=dilations[3],
                                         BatchNorm=BatchNorm)
        self._init_weight()
        if pretrained:
            self._load_pretrained_model(imagenet_pretrained_path)

    def _downsample(self, planes, stride, block, BatchNorm):
        if stride != 1 or self.inplanes != planes * block.expansion:
            return nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                BatchNorm(planes * block.expansion),
            )
        return None

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1, BatchNorm=None):
        downsample = self._downsample(planes, stride, block, BatchNorm)
        layers = [block(self.inplanes, planes, stride, dilation, downsample, BatchNorm)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, dilation=dilation, BatchNorm=BatchNorm))
        return nn.Sequential(*layers)

    def _make_MG_unit(self, block, planes, blocks, stride=1, dilation=1, BatchNorm=None):
        downsample = self._downsample(planes, stride, block, BatchNorm)
        layers = [block(self.inplanes, planes, stride, dilation=blocks[0] * dilation, downsample=downsample,
                        BatchNorm=BatchNorm)]
        self.inplanes = planes * block.expansion
        for i in range(1, len(blocks)):
            layers.append(block(self.inplanes, planes, stride=1, dilation=blocks[i] * dilation, BatchNorm=BatchNorm))
        return nn.Sequential(*layers)

    def forward(self, input):
        """input: NCHW fp32 CUDA image batch.  Returns (x, low_level_feat) as NHWC bf16 tensors."""
        if not input.is_cuda:
            raise RuntimeError("zs3_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        x = ZF.Stem.apply(self.conv1, self.bn1, self.maxpool, self.conv1.weight, self.bn1.weight, self.bn1.bias, input)
        x = self.layer1(x)
        low_level_feat = x
        x = self.layer2(x)
        x = self.layer3(x)
        x = self.layer4(x)
        return x, low_level_feat

    def _init_weight(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, (SynchronizedBatchNorm2d, nn.BatchNorm2d)):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _load_pretrained_model(self, imagenet_pretrained_path):
        # keys of the ImageNet checkpoint carry a 7-character "module." prefix (reference resnet.py:216-226)
        pretrain_dict = torch.load(imagenet_pretrained_path)["state_dict"]
        model_dict = {}
        state_dict = self.state_dict()
        for k, v in pretrain_dict.items():
            k = k[7:]
            if k in state_dict:
                model_dict[k] = v
        state_dict.update(model_dict)
        self.load_state_dict(state_dict)


def ResNet101(output_stride, BatchNorm, pretrained=True, imagenet_pretrained_path=""):
    return ResNet(Bottleneck, [3, 4, 23, 3], output_stride, BatchNorm, pretrained=pretrained,
                  imagenet_pretrained_path=imagenet_pretrained_path)
