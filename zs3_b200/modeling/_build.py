"""Construction helpers shared by the DeepLab modules: layer factories, weight initialisers and the architecture
tables.  The module *tree* (attribute names, parameter shapes) has to equal the reference's so that state_dict()s
are interchangeable; how the tree is built is table driven here."""
import math

import torch
import torch.nn as nn

# (planes, blocks) per residual stage of ResNet-101; strides / dilations depend on the output stride
RESNET101_STAGES = ((64, 3), (128, 4), (256, 23), (512, 3))
MULTI_GRID = (1, 2, 4)                      # dilation multipliers of the three layer4 blocks
STAGE_GEOMETRY = {16: ((1, 2, 2, 1), (1, 1, 1, 2)), 8: ((1, 2, 1, 1), (1, 1, 2, 4))}   # (strides, dilations)
ASPP_RATES = {16: (1, 6, 12, 18), 8: (1, 12, 24, 36)}
EXPANSION = 4


def geometry(table, output_stride):
    """row of an output-stride keyed table; unknown strides are an error exactly like in the reference"""
    if output_stride not in table:
        raise NotImplementedError(f"output_stride={output_stride}")
    return table[output_stride]


def conv(cin, cout, k=1, stride=1, dilation=1, bias=False):
    """bias-free 'same' convolution: padding = dilation * (k - 1) / 2"""
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=dilation * (k - 1) // 2, dilation=dilation,
                     bias=bias)


def is_norm(m):
    return isinstance(m, nn.modules.batchnorm._BatchNorm)


def reset_norms_(root):
    for m in root.modules():
        if is_norm(m) and m.affine:
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


def init_fan_out_(root):
    """backbone rule: N(0, sqrt(2 / (k*k*cout)))  (reference resnet.py:199-204)"""
    for m in root.modules():
        if isinstance(m, nn.Conv2d):
            fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            with torch.no_grad():
                m.weight.normal_(0.0, math.sqrt(2.0 / fan))
    reset_norms_(root)


def init_kaiming_(root):
    """head rule: kaiming_normal_ (fan_in)  (reference aspp.py:118-123, decoder.py:74-77)"""
    for m in root.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight)
    reset_norms_(root)
