from zs3_b200.modeling.sync_batchnorm import SynchronizedBatchNorm2d, patch_replication_callback  # noqa: F401
