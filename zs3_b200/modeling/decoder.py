"""DeepLabv3+ decoder on the fused kernels; mirrors zs3/modeling/decoder.py (same Sequential indices)."""
import torch
import torch.nn as nn

from .. import functional as ZF
from . import _build as B


class Decoder(nn.Module):
    """zs3/modeling/decoder.py:8-87.  Internal tensors are NHWC bf16; see DeepLab for the NCHW boundary."""

    LOW_LEVEL_WIDTH, REDUCED_WIDTH, WIDTH = 256, 48, 256
    REFINE_DROPOUT = (0.5, 0.1)     # dropout after the two 3x3 refinement convs

    def __init__(self, num_classes, BatchNorm):
        super().__init__()
        self.conv1 = B.conv(self.LOW_LEVEL_WIDTH, self.REDUCED_WIDTH, 1)
        self.bn1 = BatchNorm(self.REDUCED_WIDTH)
        self.relu = nn.ReLU()
        refine, cin = [], self.WIDTH + self.REDUCED_WIDTH
        for p in self.REFINE_DROPOUT:   # indices 0-3 and 4-7 of last_conv: conv, BN, ReLU, Dropout
            refine += [B.conv(cin, self.WIDTH, 3), BatchNorm(self.WIDTH), nn.ReLU(), nn.Dropout(p)]
            cin = self.WIDTH
        self.last_conv = nn.Sequential(*refine)
        self.pred_conv = nn.Conv2d(self.WIDTH, num_classes, kernel_size=1, stride=1)
        self.num_classes = num_classes
        B.init_kaiming_(self)

    # -- pieces -------------------------------------------------------------------------------------
    def _fuse_low_level(self, x, low_level_feat, keep_mask=None):
        """decoder.py:30-38 up to and including last_conv[0:4]"""
        low = ZF.conv_bn_act([low_level_feat], [256], self.conv1, self.bn1, relu=True)
        x = ZF.Bilinear.apply(x, low.shape[1], low.shape[2])
        lc = self.last_conv
        # torch.cat((x, low), 1) + conv: two K-segments (256 + 48 channels), concat never materialised
        return ZF.conv_bn_act([x, low], [256, 48], lc[0], lc[1], relu=True, drop_p=lc[3].p,
                              drop_training=lc[3].training, keep_mask=keep_mask)

    def _second_conv(self, x, keep_mask=None):
        lc = self.last_conv
        return ZF.conv_bn_act([x], [256], lc[4], lc[5], relu=True, drop_p=lc[7].p, drop_training=lc[7].training,
                              keep_mask=keep_mask)

    # -- reference API ------------------------------------------------------------------------------
    def forward(self, x, low_level_feat):
        x = self.forward_before_class_prediction(x, low_level_feat)
        return self.forward_class_prediction(x)

    def forward_before_class_prediction(self, x, low_level_feat, keep_masks=(None, None)):
        x = self._fuse_low_level(x, low_level_feat, keep_masks[0])
        return self._second_conv(x, keep_masks[1])

    def forward_before_last_conv_finetune(self, x, low_level_feat):
        return self._fuse_low_level(x, low_level_feat)

    def forward_class_prediction(self, x):
        """pred_conv: NHWC bf16 features (or the reference's NCHW fp32 tensor) -> NHWC bf16 class scores"""
        if x.dtype != torch.bfloat16:
            x = ZF.FromNCHW.apply(x)
        return ZF.ConvBias.apply(self.pred_conv, self.pred_conv.weight, self.pred_conv.bias, x)

    def forward_class_last_conv_finetune(self, x):
        return self._second_conv(x)

    def _init_weight(self):
        B.init_kaiming_(self)


def build_decoder(num_classes, BatchNorm):
    return Decoder(num_classes, BatchNorm)
