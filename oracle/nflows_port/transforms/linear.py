"""Linear transform base with optional caching (restating nflows.transforms.linear)."""
import torch
from torch import nn

from ..utils import torchutils
from .base import Transform


class Linear(Transform):
    def __init__(self, features, using_cache=False):
        if features <= 0:
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.features = features
        self.bias = nn.Parameter(torch.zeros(features))
        self.using_cache = using_cache

    def forward(self, inputs, context=None):
        return self.forward_no_cache(inputs)

    def inverse(self, inputs, context=None):
        return self.inverse_no_cache(inputs)

    def forward_no_cache(self, inputs):
        raise NotImplementedError()

    def inverse_no_cache(self, inputs):
        raise NotImplementedError()

    def weight(self):
        raise NotImplementedError()

    def logabsdet(self):
        raise NotImplementedError()
