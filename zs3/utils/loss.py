from zs3_b200.utils.loss import GMMNLoss, SegmentationLosses  # noqa: F401
