"""Oracle for ``CondConv.CondConvolution`` / ``CondConv.CondConvBlock``.  Test infrastructure.

The reference imports ``CondConv`` (attn_unet_data_parallel.py:28) but the module is not
in the repository, so only its call-site interface is known:

* ``CondConv.CondConvolution(dropout=0.0, is_transposed=True, num_covars=n, spatial_dims=,
  in_channels=, out_channels=, strides=, kernel_size=)``                 (:126, UpBlock)
* ``CondConv.CondConvBlock(spatial_dims=, in_channels=, out_channels=, [strides=,]
  dropout=, num_covars=5)``                                              (:289-294,318-325,360-367)
* ``CondConv.CondConvolution(spatial_dims=, in_channels=, out_channels=, kernel_size=1,
  strides=1, padding=0, conv_only=True, num_experts=8, num_covars=6)``   (:296-306)
* ``module(x, covariate)`` with ``covariate`` shaped ``[B, 1, num_covars]``  (:130,209,212,425,428)

Behaviour is SPECIFIED by this repo (parity unpinned; SURVEY.md section 7 hard part 1):

``CondConvolution`` has MONAI ``Convolution``'s signature and defaults plus ``num_experts=1``
and ``num_covars=0``.  It computes conv -> ADN where the norm output is modulated FiLM-style
per sample and channel::

    y = (1 + dgamma(c)) * norm(conv(x)) + beta(c)
    (dgamma, beta) = Linear(64 -> 2*Cout)(ReLU(Linear(num_covars -> 64)(c)))      # ``film``

with the last ``film`` layer zero-initialised, so an untrained block equals the plain MONAI
block.  With ``num_experts > 1`` the conv kernel and bias are a covariate-routed mixture
``W_b = sum_e sigmoid(Linear(num_covars -> E)(c_b))_e * W_e`` (``routing``).  Because the one
UpBlock call site passes neither ``act`` nor ``norm``, the conditional up-path uses the
``Convolution`` defaults (InstanceNorm + PReLU), not UpConv's BatchNorm + ReLU.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .monai_blocks import ADN, Norm

FILM_HIDDEN = 64


def covariate_matrix(covariate, like: torch.Tensor) -> torch.Tensor:
    """``[B,1,n]`` (float32 or float64, VolumeDataset_ADNI_A4_combined.py:86) -> ``[B,n]`` in x's dtype."""
    return covariate.reshape(covariate.shape[0], -1).to(device=like.device, dtype=like.dtype if like.is_floating_point() else torch.float32)


class ExpertConv3d(nn.Module):
    """E stacked Conv3d kernels mixed per sample by routing weights ``r[B,E]``."""

    def __init__(self, num_experts, in_channels, out_channels, kernel_size, stride, padding, bias=True):
        super().__init__()
        self.stride, self.padding = stride, padding
        self.weight = nn.Parameter(torch.empty(num_experts, out_channels, in_channels, *(kernel_size,) * 3))
        self.bias = nn.Parameter(torch.empty(num_experts, out_channels)) if bias else None
        fan_in = in_channels * kernel_size ** 3
        for e in range(num_experts):
            nn.init.kaiming_uniform_(self.weight[e], a=math.sqrt(5))
        if bias:
            nn.init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def forward(self, x, r):
        w = torch.einsum("be,eoidhw->boidhw", r, self.weight)
        b = r @ self.bias if self.bias is not None else None
        outs = [F.conv3d(x[i:i + 1], w[i], None if b is None else b[i], self.stride, self.padding)
                for i in range(x.shape[0])]
        return torch.cat(outs, dim=0)


class CondConvolution(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3,
                 adn_ordering="NDA", act="PRELU", norm="INSTANCE", dropout=None, dropout_dim=1,
                 dilation=1, groups=1, bias=True, conv_only=False, is_transposed=False,
                 padding=None, output_padding=None, num_experts=1, num_covars=0):
        super().__init__()
        assert spatial_dims == 3 and dilation == 1 and groups == 1
        self.num_experts, self.num_covars, self.ordering = num_experts, num_covars, adn_ordering.upper()
        if padding is None:
            padding = (kernel_size - 1) // 2
        if num_experts > 1:
            if is_transposed:
                raise NotImplementedError("expert-mixed transposed conv has no call site in the reference")
            self.conv = ExpertConv3d(num_experts, in_channels, out_channels, kernel_size, strides, padding, bias)
            self.routing = nn.Linear(num_covars, num_experts)
        elif is_transposed:
            if output_padding is None:
                output_padding = strides - 1
            self.conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=strides,
                                           padding=padding, output_padding=output_padding, bias=bias)
        else:
            self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=strides, padding=padding, bias=bias)
        self.adn = None
        if not conv_only and not (act is None and norm is None and dropout is None):
            self.adn = ADN(adn_ordering, out_channels, act, norm, dropout)
        self.film = None
        if self.adn is not None and hasattr(self.adn, "N") and num_covars > 0:
            self.film = nn.Sequential(nn.Linear(num_covars, FILM_HIDDEN), nn.ReLU(),
                                      nn.Linear(FILM_HIDDEN, 2 * out_channels))
            nn.init.zeros_(self.film[2].weight)
            nn.init.zeros_(self.film[2].bias)

    def forward(self, x, covariate=None):
        c = covariate_matrix(covariate, x) if covariate is not None and self.num_covars > 0 else None
        if self.num_experts > 1:
            y = self.conv(x, torch.sigmoid(self.routing(c)))
        else:
            y = self.conv(x)
        if self.adn is None:
            return y
        for key in self.ordering:
            if not hasattr(self.adn, key):
                continue
            y = getattr(self.adn, key)(y)
            if key == "N" and self.film is not None and c is not None:
                dgamma, beta = self.film(c).chunk(2, dim=-1)
                y = y * (1 + dgamma)[:, :, None, None, None] + beta[:, :, None, None, None]
        return y


class CondConvBlock(nn.Module):
    """``attentionunet.ConvBlock`` with both convolutions conditioned on the covariates."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size=3, strides=1, dropout=0.0, num_covars=0):
        super().__init__()
        common = dict(kernel_size=kernel_size, padding=None, adn_ordering="NDA", act="relu",
                      norm=Norm.BATCH, dropout=dropout, num_covars=num_covars)
        self.conv = nn.ModuleList([
            CondConvolution(spatial_dims, in_channels, out_channels, strides=strides, **common),
            CondConvolution(spatial_dims, out_channels, out_channels, strides=1, **common),
        ])

    def forward(self, x, covariate=None):
        for layer in self.conv:
            x = layer(x, covariate)
        return x
