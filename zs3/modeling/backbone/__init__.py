from zs3_b200.modeling.backbone import build_backbone, resnet  # noqa: F401
