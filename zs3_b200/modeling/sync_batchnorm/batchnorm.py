"""SynchronizedBatchNorm2d: parameter/buffer holder with the reference's class identity.

Reference: zs3/modeling/sync_batchnorm/batchnorm.py:36-142.  In the reference this class synchronises
batch statistics across torch.nn.DataParallel replica threads; on a single device it defers to
F.batch_norm (:48-58).  Here data parallelism is one process per GPU with rank-local statistics
("Multi-GPU"), so the module only has to (a) exist under this name with the same
parameters/buffers (state_dict compatibility, isinstance checks in DeepLab.get_1x_lr_params) and
(b) be consumable by the fused conv+BN kernels, which read .weight/.bias/.running_* directly.
"""
import torch.nn as nn

__all__ = ["SynchronizedBatchNorm2d"]


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
