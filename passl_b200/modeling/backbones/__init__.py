from .resnet import ResNet, ResNetsimclr  # noqa: F401
