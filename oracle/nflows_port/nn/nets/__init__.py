from .resnet import ResidualBlock, ResidualNet  # noqa: F401
