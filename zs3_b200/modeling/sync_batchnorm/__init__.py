from .batchnorm import SynchronizedBatchNorm2d
from .replicate import patch_replication_callback

__all__ = ["SynchronizedBatchNorm2d", "patch_replication_callback"]
