from . import made  # noqa: F401
from .made import MADE, MixtureOfGaussiansMADE  # noqa: F401
