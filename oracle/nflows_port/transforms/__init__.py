from .base import (  # noqa: F401
    Transform, CompositeTransform, InverseTransform, InverseNotAvailable, InputOutsideDomain,
)
from .standard import (  # noqa: F401
    IdentityTransform, PointwiseAffineTransform, AffineTransform, AffineScalarTransform,
)
from .linear import Linear  # noqa: F401
from .lu import LULinear  # noqa: F401
from .permutations import Permutation, RandomPermutation, ReversePermutation  # noqa: F401
from .coupling import (  # noqa: F401
    CouplingTransform, PiecewiseCouplingTransform, PiecewiseRationalQuadraticCouplingTransform,
)
from .autoregressive import (  # noqa: F401
    AutoregressiveTransform, MaskedAffineAutoregressiveTransform,
    MaskedPiecewiseRationalQuadraticAutoregressiveTransform,
)
from . import made, splines  # noqa: F401
