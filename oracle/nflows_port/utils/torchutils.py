"""Tensor helpers restating nflows.utils.torchutils."""
import torch

from . import typechecks as check


def tile(x, n):
    """Repeat each element of the flattened ``x`` n times consecutively: [a,a,b,b,..]."""
    if not check.is_positive_int(n):
        raise TypeError("Argument 'n' must be a positive integer.")
    x_ = x.reshape(-1)
    x_ = x_.repeat(n)
    x_ = x_.reshape(n, -1)
    x_ = x_.transpose(1, 0)
    return x_.reshape(-1)


def sum_except_batch(x, num_batch_dims=1):
    if not check.is_nonnegative_int(num_batch_dims):
        raise TypeError("Number of batch dimensions must be a non-negative integer.")
    reduce_dims = list(range(num_batch_dims, x.ndimension()))
    if not reduce_dims:
        return x
    return torch.sum(x, dim=reduce_dims)


def split_leading_dim(x, shape):
    new_shape = torch.Size(shape) + x.shape[1:]
    return torch.reshape(x, new_shape)


def merge_leading_dims(x, num_dims):
    if not check.is_positive_int(num_dims):
        raise TypeError("Number of leading dims must be a positive integer.")
    if num_dims > x.dim():
        raise ValueError("Number of leading dims can't be greater than total dims.")
    new_shape = torch.Size([-1]) + x.shape[num_dims:]
    return torch.reshape(x, new_shape)


def repeat_rows(x, num_reps):
    """Each row repeated num_reps times, consecutively (row-wise interleave)."""
    if not check.is_positive_int(num_reps):
        raise TypeError("Number of repetitions must be a positive integer.")
    shape = x.shape
    x = x.unsqueeze(1)
    x = x.expand(shape[0], num_reps, *shape[1:])
    return merge_leading_dims(x, num_dims=2)


def searchsorted(bin_locations, inputs, eps=1e-6):
    """bin = #(knots <= x) - 1, the last knot nudged up by eps (IN PLACE, as nflows)."""
    bin_locations[..., -1] += eps
    return torch.sum(inputs[..., None] >= bin_locations, dim=-1) - 1


def get_num_parameters(model):
    return sum(p.numel() for p in model.parameters())
