"""DeepLabv3+ top module with the reference's API (zs3/modeling/deeplab.py:10-99).

Inputs/outputs at this boundary are the reference's: NCHW fp32 CUDA tensors that take part in autograd.
Inside, activations are NHWC bf16 and every op is one of the hand-written sm_100a kernels behind
include/zs3b200.h.
"""
import torch.nn as nn

from .. import functional as ZF
from . import _build as B
from .aspp import build_aspp
from .backbone import build_backbone
from .decoder import build_decoder
from .sync_batchnorm.batchnorm import SynchronizedBatchNorm2d


class DeepLab(nn.Module):
    def __init__(self, output_stride=16, num_classes=21, sync_bn=True, freeze_bn=False, pretrained=True,
                 global_avg_pool_bn=True, imagenet_pretrained_path=""):
        super().__init__()
        BatchNorm = SynchronizedBatchNorm2d if sync_bn else nn.BatchNorm2d
        self.backbone = build_backbone(output_stride, BatchNorm, pretrained=pretrained,
                                       imagenet_pretrained_path=imagenet_pretrained_path)
        self.aspp = build_aspp(output_stride, BatchNorm, global_avg_pool_bn)
        self.decoder = build_decoder(num_classes, BatchNorm)
        self.num_classes = num_classes
        # conv parameters live in HBM in KRSC order (torch channels_last): same names/shapes/values as the
        # reference's OIHW tensors, but directly consumable by the packing and weight-gradient kernels
        ZF.to_krsc_(self)
        if freeze_bn:
            self.freeze_bn()

    # ---------------------------------------------------------------------------------------- forward
    def _features(self, input, keep_masks=None):
        km = keep_masks or {}
        if input.is_cuda:
            ZF.reset_stat_buffers(input.device)
        x, low_level_feat = self.backbone(input)
        x = self.aspp(x, km.get("aspp.dropout"))
        x = self.decoder.forward_before_class_prediction(
            x, low_level_feat, (km.get("decoder.dropout0"), km.get("decoder.dropout1")))
        ZF.flush_batch_counters()
        return x

    def forward(self, input, keep_masks=None):
        """deeplab.py:40-45.  keep_masks: optional {"aspp.dropout","decoder.dropout0","decoder.dropout1"} ->
        uint8 NHWC keep masks (parity tests inject the oracle's masks; None = counter-based RNG)."""
        x = self._features(input, keep_masks)
        x = self.decoder.forward_class_prediction(x)
        return ZF.UpsampleLogits.apply(x, self.num_classes, input.shape[2], input.shape[3])

    def forward_scores(self, input, keep_masks=None):
        """Class scores BEFORE the final x4 upsample: NHWC bf16 [N, H/4, W/4, cpad(num_classes)].  The training
        runtime feeds them to SegmentationLosses.UpsampledCrossEntropyLoss, which fuses the upsample into the loss."""
        return self.decoder.forward_class_prediction(self._features(input, keep_masks))

    def forward_before_class_prediction(self, input, keep_masks=None):
        """deeplab.py:47-51: returns the [B,256,H/4,W/4] decoder features as an NCHW fp32 tensor"""
        return ZF.ToNCHW.apply(self._features(input, keep_masks), 256)

    def forward_class_prediction(self, x, input_size):
        """deeplab.py:53-56: x is an NCHW fp32 feature tensor (real or generated features)"""
        x = self.decoder.forward_class_prediction(x)
        return ZF.UpsampleLogits.apply(x, self.num_classes, int(input_size[0]), int(input_size[1]))

    def forward_before_last_conv_finetune(self, input):
        if input.is_cuda:
            ZF.reset_stat_buffers(input.device)
        x, low_level_feat = self.backbone(input)
        x = self.aspp(x)
        x = self.decoder.forward_before_last_conv_finetune(x, low_level_feat)
        ZF.flush_batch_counters()
        return ZF.ToNCHW.apply(x, 256)

    def forward_class_last_conv_finetune(self, x):
        x = self.decoder.forward_class_last_conv_finetune(ZF.FromNCHW.apply(x))
        ZF.flush_batch_counters()
        return ZF.ToNCHW.apply(x, 256)

    # ------------------------------------------------------------------------------- reference helpers
    def freeze_bn(self):
        for m in filter(B.is_norm, self.modules()):
            m.eval()

    @staticmethod
    def _trainable(*roots):
        """parameters of the conv / norm layers under `roots` that still require grad (deeplab.py:71-99)"""
        for root in roots:
            for m in root.modules():
                if isinstance(m, nn.Conv2d) or B.is_norm(m):
                    yield from (p for p in m.parameters(recurse=False) if p.requires_grad)

    def get_1x_lr_params(self):
        return self._trainable(self.backbone)

    def get_10x_lr_params(self):
        return self._trainable(self.aspp, self.decoder)
