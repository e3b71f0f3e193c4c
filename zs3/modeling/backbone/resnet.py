from zs3_b200.modeling.backbone.resnet import Bottleneck, ResNet, ResNet101  # noqa: F401
