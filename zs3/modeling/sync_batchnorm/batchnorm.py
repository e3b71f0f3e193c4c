from zs3_b200.modeling.sync_batchnorm.batchnorm import SynchronizedBatchNorm2d  # noqa: F401
