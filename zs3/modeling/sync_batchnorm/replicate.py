from zs3_b200.modeling.sync_batchnorm.replicate import patch_replication_callback  # noqa: F401
