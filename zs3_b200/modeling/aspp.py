"""ASPP (atrous spatial pyramid pooling) on the fused kernels; parameter tree as in zs3/modeling/aspp.py."""
import torch.nn as nn

from .. import functional as ZF
from . import _build as B

ASPP_WIDTH = 256
BACKBONE_WIDTH = 2048


class _ASPPModule(nn.Module):
    """one pyramid branch: (atrous) conv -> BN -> ReLU, reference aspp.py:8-40"""

    def __init__(self, inplanes, planes, kernel_size, padding, dilation, BatchNorm):
        super().__init__()
        self.inplanes = inplanes
        self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, stride=1, padding=padding,
                                     dilation=dilation, bias=False)
        self.bn = BatchNorm(planes)
        self.relu = nn.ReLU()
        B.init_kaiming_(self)

    def forward(self, x):
        return ZF.conv_bn_act([x], [self.inplanes], self.atrous_conv, self.bn, relu=True)


class ASPP(nn.Module):
    """four atrous branches + image-level pooling branch -> 1x1 projection -> Dropout, reference aspp.py:43-116"""

    def __init__(self, output_stride, BatchNorm, global_avg_pool_bn=True):
        super().__init__()
        rates = B.geometry(B.ASPP_RATES, output_stride)
        self.inplanes = BACKBONE_WIDTH
        for i, rate in enumerate(rates, start=1):
            k = 1 if i == 1 else 3
            setattr(self, f"aspp{i}", _ASPPModule(BACKBONE_WIDTH, ASPP_WIDTH, k, 0 if k == 1 else rate, rate, BatchNorm))
        pool = [nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(BACKBONE_WIDTH, ASPP_WIDTH, 1, stride=1, bias=False)]
        if global_avg_pool_bn:
            pool.append(BatchNorm(ASPP_WIDTH))
        self.global_avg_pool = nn.Sequential(*pool, nn.ReLU())
        self.global_avg_pool_bn = global_avg_pool_bn
        self.conv1 = nn.Conv2d(5 * ASPP_WIDTH, ASPP_WIDTH, 1, bias=False)
        self.bn1 = BatchNorm(ASPP_WIDTH)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)
        B.init_kaiming_(self)

    def forward(self, x, keep_mask=None):
        """x: NHWC bf16 backbone features.  One fused autograd node: four atrous branches + image-level branch (global
        mean -> 1x1 conv (-> BN) -> ReLU -> broadcast, the bilinear "upsampling" from 1x1) -> concat + 1x1 projection
        as five K-segments of one implicit GEMM -> BN -> ReLU -> Dropout."""
        return ZF.aspp_head(self, x, keep_mask)


def build_aspp(output_stride, BatchNorm, global_avg_pool_bn=True):
    return ASPP(output_stride, BatchNorm, global_avg_pool_bn)
