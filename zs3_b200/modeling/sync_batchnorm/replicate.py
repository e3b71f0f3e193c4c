"""patch_replication_callback kept for import compatibility with the reference trainers
(zs3/train_pascal_GMMN.py:12,96-97; reference implementation zs3/modeling/sync_batchnorm/replicate.py:45-68).

The reference patches DataParallel.replicate so its SyncBN replicas can find each other.  This build runs
one process per GPU (torch.distributed + NCCL, zs3_b200/parallel.py), a DataParallel wrapper therefore
always has exactly one replica and there is nothing to patch: the function validates its argument and
returns.
"""
from torch.nn.parallel.data_parallel import DataParallel

__all__ = ["patch_replication_callback"]


def patch_replication_callback(data_parallel):
    assert isinstance(data_parallel, DataParallel)
    return None
