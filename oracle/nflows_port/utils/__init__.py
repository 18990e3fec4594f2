from . import torchutils, typechecks  # noqa: F401
from .torchutils import (  # noqa: F401
    tile, sum_except_batch, split_leading_dim, merge_leading_dims, repeat_rows,
    searchsorted, get_num_parameters,
)
