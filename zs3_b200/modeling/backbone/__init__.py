"""Backbones of the DeepLab encoder; `build_backbone` is the factory the top module calls."""
from .resnet import BACKBONES, build_backbone  # noqa: F401
