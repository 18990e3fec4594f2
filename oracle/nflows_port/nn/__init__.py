from . import nets, nde  # noqa: F401
