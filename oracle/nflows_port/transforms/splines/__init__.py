from . import rational_quadratic  # noqa: F401
from .rational_quadratic import (  # noqa: F401
    rational_quadratic_spline, unconstrained_rational_quadratic_spline,
)
