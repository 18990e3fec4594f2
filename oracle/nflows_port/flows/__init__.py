from .base import Flow  # noqa: F401
