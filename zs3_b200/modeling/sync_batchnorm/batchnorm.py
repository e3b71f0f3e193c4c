"""SynchronizedBatchNorm2d: the reference's class identity plus cross-rank synchronisation of the batch statistics.

Reference: zs3/modeling/sync_batchnorm/batchnorm.py:36-142.  There the class synchronises batch statistics across
torch.nn.DataParallel replica threads (per layer: ReduceAddCoalesced of (sum, sum of squares) to a master replica,
`_compute_mean_std`, Broadcast of (mean, inv_std) back, plus the mirrored reduction autograd adds in the backward);
on a single device it defers to F.batch_norm (:48-58).

Here data parallelism is one process per GPU.  The module itself only holds parameters and buffers (state_dict
compatibility, isinstance checks in DeepLab.get_1x_lr_params) -- the arithmetic lives in the fused kernels.  When
synchronisation is switched on (`enable_sync(world_size)`, done by DataParallelTrainer(sync_bn=True)):

  * forward: the per-channel (sum, sum of squares) the conv epilogue produced on every rank is all-reduced (one
    32 KB fp64 NCCL message per layer, no master replica, no broadcast step) between the conv kernel and the
    normalisation kernel, whose fused finalize then uses the GLOBAL count and the reference's multi-replica
    formula inv_std = max(var, eps)^-1/2 (`_compute_mean_std`, :124-142; note: clamp, not var + eps); every rank
    updates its (identical) running statistics with the unbiased global variance;
  * backward: the two per-channel sums of the BatchNorm backward (sum dz, sum dz * xhat) are all-reduced between the
    reduction kernel and the apply kernel (what autograd derives from ReduceAddCoalesced / Broadcast in the
    reference); dgamma / dbeta keep the rank-local sums, they join the step's gradient all-reduce.

Default is OFF: rank-local statistics (bs = 16 per GPU), the measured configuration of SCALE_r*.json.
"""
import torch.nn as nn

__all__ = ["SynchronizedBatchNorm2d", "enable_sync", "sync_world"]

_SYNC = {"world": 1, "group": None}


def enable_sync(world_size, group=None):
    """world_size > 1: synchronise training-mode statistics of every SynchronizedBatchNorm2d over the ranks of
    `group` (default process group); 1 switches it off.  Eager launches only (the collectives sit between kernels)."""
    _SYNC["world"], _SYNC["group"] = max(1, int(world_size)), group


def sync_world(bn):
    """number of ranks whose statistics `bn` combines in training mode (1 = rank-local)"""
    return _SYNC["world"] if isinstance(bn, SynchronizedBatchNorm2d) else 1


def all_reduce_stats(t):
    import torch.distributed as dist
    dist.all_reduce(t, group=_SYNC["group"])


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
