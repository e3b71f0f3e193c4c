"""zs3_b200: B200-native (sm_100a) implementation of the ZS3Net training hot path.

Public surface mirrors the reference package (valeoai/ZS3):
    zs3_b200.modeling.deeplab.DeepLab, zs3_b200.modeling.gmmn.GMMNnetwork(_GCN),
    zs3_b200.utils.loss.{SegmentationLosses, GMMNLoss},
    zs3_b200.modeling.sync_batchnorm.{SynchronizedBatchNorm2d, patch_replication_callback}
and is re-exported under the reference's own import paths by the `zs3` shim package at the repo root.
"""
__version__ = "0.1.0"
