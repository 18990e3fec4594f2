"""Transform base classes (restating nflows.transforms.base)."""
import torch
from torch import nn


class InverseNotAvailable(Exception):
    pass


class InputOutsideDomain(Exception):
    pass


class Transform(nn.Module):
    """Invertible map returning (outputs, log|det J|)."""

    def forward(self, inputs, context=None):
        raise NotImplementedError()

    def inverse(self, inputs, context=None):
        raise InverseNotAvailable()


class CompositeTransform(Transform):
    def __init__(self, transforms):
        super().__init__()
        self._transforms = nn.ModuleList(transforms)

    @staticmethod
    def _cascade(inputs, funcs, context):
        batch_size = inputs.shape[0]
        outputs = inputs
        total_logabsdet = inputs.new_zeros(batch_size)
        for func in funcs:
            outputs, logabsdet = func(outputs, context)
            total_logabsdet = total_logabsdet + logabsdet
        return outputs, total_logabsdet

    def forward(self, inputs, context=None):
        return self._cascade(inputs, self._transforms, context)

    def inverse(self, inputs, context=None):
        funcs = (t.inverse for t in self._transforms[::-1])
        return self._cascade(inputs, funcs, context)


class InverseTransform(Transform):
    def __init__(self, transform):
        super().__init__()
        self._transform = transform

    def forward(self, inputs, context=None):
        return self._transform.inverse(inputs, context)

    def inverse(self, inputs, context=None):
        return self._transform(inputs, context)
