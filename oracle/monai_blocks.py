"""Oracle restatement of the MONAI blocks the reference subclasses.  Test infrastructure.

The arithmetic of the hot path lives in ``monai`` (unpinned third-party dependency of
the reference, absent from /root/reference and from this image).  Call sites that fix
the interface: attn_unet_data_parallel.py:20-23 (imports), :120,134,152 (subclassing
``attentionunet.UpConv/AttentionBlock/AttentionLayer``), :285-286,442,495-497,547-548,558
(``ConvBlock`` / ``Convolution``).  Restated from MONAI's published
``monai/networks/blocks/{convolutions,acti_norm}.py`` and
``monai/networks/nets/attentionunet.py``; parity unpinned.

Child-module names (``conv``, ``adn``, ``N``/``D``/``A``, ``W_g``/``W_x``/``psi``, ``up``,
``attention``/``upconv``/``merge``/``submodule``) follow MONAI so ``state_dict`` keys match
checkpoints written by the reference (attn_unet_data_parallel.py:946-953).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class Norm:  # monai.networks.layers.factories.Norm (only the members the reference touches)
    BATCH = "BATCH"
    INSTANCE = "INSTANCE"


def make_norm(spec, channels: int) -> nn.Module:
    name = spec[0] if isinstance(spec, (tuple, list)) else spec
    name = str(name).upper()
    if name == "BATCH":
        return nn.BatchNorm3d(channels)  # affine, eps 1e-5, momentum 0.1
    if name == "INSTANCE":
        return nn.InstanceNorm3d(channels)  # no affine, no running stats, eps 1e-5
    raise ValueError(f"norm {spec!r} not used on the hot path")


def make_act(spec) -> nn.Module:
    kwargs = {}
    if isinstance(spec, (tuple, list)):
        spec, kwargs = spec[0], dict(spec[1])
    if isinstance(spec, type):  # MONAI's factory returns non-string keys unchanged
        return spec(**kwargs)
    name = str(spec).upper()
    if name == "PRELU":
        return nn.PReLU(**kwargs)  # one parameter, init 0.25
    if name == "RELU":
        return nn.ReLU(**kwargs)
    if name == "LEAKYRELU":
        return nn.LeakyReLU(**kwargs)
    raise ValueError(f"activation {spec!r} not used on the hot path")


class ADN(nn.Sequential):
    """Activation / Dropout / Norm in ``ordering`` order; a member is present iff its spec is not None."""

    def __init__(self, ordering="NDA", in_channels=None, act="RELU", norm=None, dropout=None):
        super().__init__()
        parts = {"A": None, "D": None, "N": None}
        if norm is not None:
            parts["N"] = make_norm(norm, in_channels)
        if act is not None:
            parts["A"] = make_act(act)
        if dropout is not None:
            parts["D"] = nn.Dropout(p=float(dropout))  # dropout=0.0 still adds Dropout(p=0)
        for key in ordering.upper():
            if parts[key] is not None:
                self.add_module(key, parts[key])


class Convolution(nn.Sequential):
    """``conv`` (+ ``adn``): Conv3d / ConvTranspose3d, 'same' padding, output_padding = stride-1."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3,
                 adn_ordering="NDA", act="PRELU", norm="INSTANCE", dropout=None, dropout_dim=1,
                 dilation=1, groups=1, bias=True, conv_only=False, is_transposed=False,
                 padding=None, output_padding=None):
        super().__init__()
        assert spatial_dims == 3 and dilation == 1 and groups == 1
        self.spatial_dims, self.in_channels, self.out_channels = spatial_dims, in_channels, out_channels
        self.is_transposed = is_transposed
        if padding is None:
            padding = (kernel_size - 1) // 2
        if is_transposed:
            if output_padding is None:
                output_padding = strides - 1
            conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=strides,
                                      padding=padding, output_padding=output_padding, bias=bias)
        else:
            conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=strides, padding=padding, bias=bias)
        self.add_module("conv", conv)
        if conv_only or (act is None and norm is None and dropout is None):
            return
        self.add_module("adn", ADN(adn_ordering, out_channels, act, norm, dropout))


class ConvBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size=3, strides=1, dropout=0.0):
        super().__init__()
        common = dict(kernel_size=kernel_size, padding=None, adn_ordering="NDA", act="relu",
                      norm=Norm.BATCH, dropout=dropout)
        self.conv = nn.Sequential(
            Convolution(spatial_dims, in_channels, out_channels, strides=strides, **common),
            Convolution(spatial_dims, out_channels, out_channels, strides=1, **common),
        )

    def forward(self, x):
        return self.conv(x)


class UpConv(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size=3, strides=2, dropout=0.0):
        super().__init__()
        self.up = Convolution(spatial_dims, in_channels, out_channels, strides=strides, kernel_size=kernel_size,
                              act="relu", adn_ordering="NDA", norm=Norm.BATCH, dropout=dropout, is_transposed=True)

    def forward(self, x):
        return self.up(x)


class AttentionBlock(nn.Module):
    def __init__(self, spatial_dims, f_int, f_g, f_l, dropout=0.0):
        super().__init__()

        def pointwise(cin, cout):
            return Convolution(spatial_dims, cin, cout, kernel_size=1, strides=1, padding=0,
                               dropout=dropout, conv_only=True)

        self.W_g = nn.Sequential(pointwise(f_g, f_int), nn.BatchNorm3d(f_int))
        self.W_x = nn.Sequential(pointwise(f_l, f_int), nn.BatchNorm3d(f_int))
        self.psi = nn.Sequential(pointwise(f_int, 1), nn.BatchNorm3d(1), nn.Sigmoid())
        self.relu = nn.ReLU()

    def forward(self, g, x):
        return x * self.psi(self.relu(self.W_g(g) + self.W_x(x)))


class AttentionLayer(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, submodule, up_kernel_size=3, strides=2, dropout=0.0):
        super().__init__()
        self.attention = AttentionBlock(spatial_dims, f_g=in_channels, f_l=in_channels, f_int=in_channels // 2)
        self.upconv = UpConv(spatial_dims, out_channels, in_channels, strides=strides, kernel_size=up_kernel_size)
        self.merge = Convolution(spatial_dims, 2 * in_channels, in_channels, dropout=dropout)
        self.submodule = submodule

    def forward(self, x):
        fromlower = self.upconv(self.submodule(x))
        att = self.attention(g=fromlower, x=x)
        return self.merge(torch.cat((att, fromlower), dim=1))
