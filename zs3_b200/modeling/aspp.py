"""ASPP (atrous spatial pyramid pooling) on the fused kernels; mirrors zs3/modeling/aspp.py."""
import torch
import torch.nn as nn

from .. import functional as ZF
from .sync_batchnorm.batchnorm import SynchronizedBatchNorm2d


def _init(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            torch.nn.init.kaiming_normal_(m.weight)
        elif isinstance(m, (SynchronizedBatchNorm2d, nn.BatchNorm2d)):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


class _ASPPModule(nn.Module):
    """zs3/modeling/aspp.py:8-40"""

    def __init__(self, inplanes, planes, kernel_size, padding, dilation, BatchNorm):
        super().__init__()
        self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, stride=1, padding=padding,
                                     dilation=dilation, bias=False)
        self.bn = BatchNorm(planes)
        self.relu = nn.ReLU()
        self.inplanes = inplanes
        _init(self)

    def forward(self, x):
        return ZF.conv_bn_act([x], [self.inplanes], self.atrous_conv, self.bn, relu=True)


class ASPP(nn.Module):
    """zs3/modeling/aspp.py:43-116"""

    def __init__(self, output_stride, BatchNorm, global_avg_pool_bn=True):
        super().__init__()
        inplanes = 2048
        if output_stride == 16:
            dilations = [1, 6, 12, 18]
        elif output_stride == 8:
            dilations = [1, 12, 24, 36]
        else:
            raise NotImplementedError
        self.aspp1 = _ASPPModule(inplanes, 256, 1, padding=0, dilation=dilations[0], BatchNorm=BatchNorm)
        self.aspp2 = _ASPPModule(inplanes, 256, 3, padding=dilations[1], dilation=dilations[1], BatchNorm=BatchNorm)
        self.aspp3 = _ASPPModule(inplanes, 256, 3, padding=dilations[2], dilation=dilations[2], BatchNorm=BatchNorm)
        self.aspp4 = _ASPPModule(inplanes, 256, 3, padding=dilations[3], dilation=dilations[3], BatchNorm=BatchNorm)
        if global_avg_pool_bn:
            self.global_avg_pool = nn.Sequential(
                nn.AdaptiveAvgPool2d((1, 1)),
                nn.Conv2d(inplanes, 256, 1, stride=1, bias=False),
                BatchNorm(256),
                nn.ReLU(),
            )
        else:
            self.global_avg_pool = nn.Sequential(
                nn.AdaptiveAvgPool2d((1, 1)),
                nn.Conv2d(inplanes, 256, 1, stride=1, bias=False),
                nn.ReLU(),
            )
        self.global_avg_pool_bn = global_avg_pool_bn
        self.conv1 = nn.Conv2d(1280, 256, 1, bias=False)
        self.bn1 = BatchNorm(256)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)
        self.inplanes = inplanes
        _init(self)

    def forward(self, x, keep_mask=None):
        n, h, w, _ = x.shape
        x1 = self.aspp1(x)
        x2 = self.aspp2(x)
        x3 = self.aspp3(x)
        x4 = self.aspp4(x)
        g = ZF.SpatialMean.apply(x)
        gconv = self.global_avg_pool[1]
        gbn = self.global_avg_pool[2] if self.global_avg_pool_bn else ZF.IdentityBN(256, x.device)
        g = ZF.conv_bn_act([g], [self.inplanes], gconv, gbn, relu=True)
        # F.interpolate of a 1x1 map with align_corners=True is a pure broadcast (aspp.py:109)
        x5 = ZF.SpatialBroadcast.apply(g, h, w)
        # torch.cat + conv1 (aspp.py:110-112): the concat is never materialised, conv1 reduces over 5 K-segments
        return ZF.conv_bn_act([x1, x2, x3, x4, x5], [256] * 5, self.conv1, self.bn1, relu=True,
                              drop_p=self.dropout.p, drop_training=self.dropout.training, keep_mask=keep_mask)


def build_aspp(output_stride, BatchNorm, global_avg_pool_bn=True):
    return ASPP(output_stride, BatchNorm, global_avg_pool_bn)
