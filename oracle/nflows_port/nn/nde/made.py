"""MADE density estimators (restating nflows.nn.nde.made).

The masked network itself is shared with ``transforms.made``; this module adds the
mixture-of-Gaussians head that sbi's ``build_made`` uses through ``MADEMoGWrapper``
(sbi/utils/nn_utils.py:133-201)."""
import numpy as np
import torch
from torch.nn import functional as F

from ...transforms.made import (  # noqa: F401
    MADE, MaskedFeedforwardBlock, MaskedLinear, MaskedResidualBlock, _get_input_degrees,
)


class MixtureOfGaussiansMADE(MADE):
    def __init__(self, features, hidden_features, context_features=None, num_blocks=2,
                 num_mixture_components=1, use_residual_blocks=True, random_mask=False,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False,
                 epsilon=1e-2, custom_initialization=True):
        if use_residual_blocks and random_mask:
            raise ValueError("Residual blocks can't be used with random masks.")
        super().__init__(
            features, hidden_features, context_features=context_features,
            num_blocks=num_blocks, output_multiplier=3 * num_mixture_components,
            use_residual_blocks=use_residual_blocks, random_mask=random_mask,
            activation=activation, dropout_probability=dropout_probability,
            use_batch_norm=use_batch_norm,
        )
        self.num_mixture_components = num_mixture_components
        self.features = features
        self.hidden_features = hidden_features
        self.epsilon = epsilon
        if custom_initialization:
            self._initialize()

    def forward(self, inputs, context=None):
        return super().forward(inputs, context=context)

    def log_prob(self, inputs, context=None):
        outputs = self.forward(inputs, context=context)
        outputs = outputs.reshape(*inputs.shape, self.num_mixture_components, 3)
        logits, means, unconstrained_stds = (
            outputs[..., 0], outputs[..., 1], outputs[..., 2],
        )
        log_mixture_coefficients = torch.log_softmax(logits, dim=-1)
        stds = F.softplus(unconstrained_stds) + self.epsilon
        log_prob = torch.sum(
            torch.logsumexp(
                log_mixture_coefficients
                - 0.5 * (
                    np.log(2 * np.pi)
                    + 2 * torch.log(stds)
                    + ((inputs[..., None] - means) / stds) ** 2
                ),
                dim=-1,
            ),
            dim=-1,
        )
        return log_prob

    def sample(self, num_samples, context=None):
        if context is not None:
            context = torch.repeat_interleave(context, num_samples, dim=0)
        with torch.no_grad():
            samples = torch.zeros(context.shape[0], self.features)
            for feature in range(self.features):
                outputs = self.forward(samples, context)
                outputs = outputs.reshape(*samples.shape, self.num_mixture_components, 3)
                logits, means, unconstrained_stds = (
                    outputs[:, feature, :, 0], outputs[:, feature, :, 1],
                    outputs[:, feature, :, 2],
                )
                logits = torch.log_softmax(logits, dim=-1)
                stds = F.softplus(unconstrained_stds) + self.epsilon
                component_distribution = torch.distributions.Categorical(logits=logits)
                components = component_distribution.sample((1,)).reshape(-1, 1)
                means, stds = (
                    means.gather(1, components).reshape(-1),
                    stds.gather(1, components).reshape(-1),
                )
                samples[:, feature] = (means + torch.randn(context.shape[0]) * stds).detach()
        return samples.reshape(-1, num_samples, self.features)

    def _initialize(self):
        # logits ~ small noise, means ~ N(0, eps-ish), std such that softplus(.)+eps = 1
        self.final_layer.weight.data[::3, :] = self.epsilon * torch.randn(
            self.features * self.num_mixture_components, self.hidden_features
        )
        self.final_layer.bias.data[::3] = self.epsilon * torch.randn(
            self.features * self.num_mixture_components
        )
        self.final_layer.weight.data[2::3] = self.epsilon * torch.randn(
            self.features * self.num_mixture_components, self.hidden_features
        )
        self.final_layer.bias.data[2::3] = torch.log(
            torch.exp(torch.Tensor([1 - self.epsilon])) - 1
        ) * torch.ones(
            self.features * self.num_mixture_components
        ) + self.epsilon * torch.randn(self.features * self.num_mixture_components)
