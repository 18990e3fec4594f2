from .base import Distribution, NoMeanException  # noqa: F401
from .normal import StandardNormal  # noqa: F401
from .mixture import MADEMoG  # noqa: F401
