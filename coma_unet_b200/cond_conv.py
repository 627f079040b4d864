"""``CondConv.CondConvolution`` / ``CondConv.CondConvBlock`` on the CUDA kernels.

The reference imports ``CondConv`` (attn_unet_data_parallel.py:28) but does not ship it; the call
sites (:126, :285-306, :318-325, :354-367; forward calls :130,209,212,425,428) fix the interface and
DESIGN.md fixes the behaviour (same specification as oracle/cond_conv.py):

    y = act((1 + dgamma(c)) * norm(conv(x)) + beta(c)),   (dgamma, beta) = film(c)
    num_experts > 1:  W_b = sum_e sigmoid(routing(c_b))_e * W_e   (per-sample kernel and bias)

The FiLM MLP and the expert mixing are O(B*C) vector math (torch); the modulation itself is applied
inside the fused normalise+modulate+activate kernel (ops.norm_act) or the conv epilogue.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .blocks import Convolution, Norm

FILM_HIDDEN = 64


def covariate_matrix(covariate, like):
    """``[B,1,n]`` float32/float64 (VolumeDataset_ADNI_A4_combined.py:86) -> ``[B,n]`` float32 on x's device."""
    if covariate.device == like.device and covariate.dtype == torch.float32:
        return covariate.reshape(covariate.shape[0], -1)      # the model moved/cast it once per forward
    return covariate.reshape(covariate.shape[0], -1).to(device=like.device, dtype=torch.float32)


class FilmBatch:
    """All FiLM MLPs of a model evaluated in two batched GEMMs (no-grad paths).

    Per conditioned layer the reference's module would run Linear -> ReLU -> Linear on a ``[B, num_covars]`` matrix: ~8 tiny
    kernels x ~14 layers on the critical stream.  The MLPs only depend on the covariates, so ``compute`` evaluates them
    all at once from stacked, zero-padded weights (columns beyond a layer's ``num_covars`` and rows beyond its ``2*Cout``
    are zero) and ``CondConvolution.forward`` picks its ``(dgamma, beta)`` slice up from ``active``.
    """

    active: dict = {}

    def __init__(self, model):
        self.mods = [m for m in model.modules() if isinstance(m, CondConvolution) and m.film is not None]
        self._key, self._w = None, None

    def _stacked(self, device):
        params = [p for m in self.mods for p in m.film.parameters()]
        key = (str(device), ops.weight_epoch()) + tuple((p._version, p.data_ptr()) for p in params)
        if key != self._key:
            L = len(self.mods)
            nmax = max(m.num_covars for m in self.mods)
            cmax = max(m.out_channels for m in self.mods)
            W1 = torch.zeros(L, nmax, FILM_HIDDEN, device=device)
            b1 = torch.zeros(L, 1, FILM_HIDDEN, device=device)
            W2 = torch.zeros(L, FILM_HIDDEN, 2 * cmax, device=device)
            b2 = torch.zeros(L, 1, 2 * cmax, device=device)
            with torch.no_grad():
                for l, m in enumerate(self.mods):
                    c = m.out_channels
                    W1[l, :m.num_covars] = m.film[0].weight.t()
                    b1[l, 0] = m.film[0].bias
                    W2[l, :, :c] = m.film[2].weight[:c].t()
                    W2[l, :, cmax:cmax + c] = m.film[2].weight[c:].t()
                    b2[l, 0, :c] = m.film[2].bias[:c]
                    b2[l, 0, cmax:cmax + c] = m.film[2].bias[c:]
            self._key, self._w = key, (W1, b1, W2, b2, nmax, cmax)
        return self._w

    def compute(self, covariate):
        """covariate: ``[B,1,n]`` / ``[B,n]`` float32 on the device.  Fills ``FilmBatch.active`` for the coming forward."""
        if not self.mods:
            return
        W1, b1, W2, b2, nmax, cmax = self._stacked(covariate.device)
        c = covariate.reshape(covariate.shape[0], -1)[:, :nmax]
        L = len(self.mods)
        hid = torch.relu(torch.baddbmm(b1, c.unsqueeze(0).expand(L, -1, -1), W1))        # [L, B, 64]
        out = torch.baddbmm(b2, hid, W2)                                                  # [L, B, 2*cmax]
        FilmBatch.active = {id(m): (out[l, :, :m.out_channels], out[l, :, cmax:cmax + m.out_channels])
                            for l, m in enumerate(self.mods)}

    def compute_fused(self, covariate):
        """The same slices from ONE fused launch (ops.FilmAllFn; one more in backward) instead of a Linear -> ReLU -> Linear chain
        per layer under autograd (~185 framework launches per training step of the north-star model) or two batched GEMMs plus
        a strided-slice copy per layer without it.  False = not applicable."""
        c = covariate.reshape(covariate.shape[0], -1)
        if not (ops.FUSED_FILM and self.mods and c.is_cuda and c.dtype == torch.float32 and c.shape[0] <= 64
                and len(self.mods) <= 32 and all(p.dtype == torch.float32 for m in self.mods for p in m.film.parameters())):
            return False
        if c.stride(1) != 1:
            c = c.contiguous()
        meta = tuple((m.num_covars, m.out_channels) for m in self.mods)
        params = [p for m in self.mods for p in (m.film[0].weight, m.film[0].bias, m.film[2].weight, m.film[2].bias)]
        outs = ops.FilmAllFn.apply(c, meta, *params)
        FilmBatch.active = {id(m): (outs[2 * l], outs[2 * l + 1]) for l, m in enumerate(self.mods)}
        return True

    @staticmethod
    def clear():
        FilmBatch.active = {}


class ExpertConv3d(nn.Module):
    """Parameter container: E stacked conv kernels ``[E,Cout,Cin,k,k,k]`` and biases ``[E,Cout]``."""

    def __init__(self, num_experts, in_channels, out_channels, kernel_size, stride, padding, bias=True):
        super().__init__()
        self.stride, self.padding, self.kernel_size = stride, padding, kernel_size
        self.weight = nn.Parameter(torch.empty(num_experts, out_channels, in_channels, *(kernel_size,) * 3))
        self.bias = nn.Parameter(torch.empty(num_experts, out_channels)) if bias else None
        fan_in = in_channels * kernel_size ** 3
        for e in range(num_experts):
            nn.init.kaiming_uniform_(self.weight[e], a=math.sqrt(5))
        if bias:
            nn.init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))


class CondConvolution(Convolution):
    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, adn_ordering="NDA",
                 act="PRELU", norm="INSTANCE", dropout=None, dropout_dim=1, dilation=1, groups=1, bias=True,
                 conv_only=False, is_transposed=False, padding=None, output_padding=None, num_experts=1, num_covars=0):
        if num_experts > 1:
            if is_transposed or kernel_size != 1 or strides != 1 or not conv_only:
                raise NotImplementedError("expert mixing is only used on the 1x1x1 conv_only reduce_channels "
                                          "(attn_unet_data_parallel.py:296-306)")
            nn.Module.__init__(self)
            self.in_channels, self.out_channels = in_channels, out_channels
            self.conv = ExpertConv3d(num_experts, in_channels, out_channels, kernel_size, strides, 0, bias)
            self.routing = nn.Linear(num_covars, num_experts)
        else:
            super().__init__(spatial_dims, in_channels, out_channels, strides=strides, kernel_size=kernel_size,
                             adn_ordering=adn_ordering, act=act, norm=norm, dropout=dropout, dilation=dilation,
                             groups=groups, bias=bias, conv_only=conv_only, is_transposed=is_transposed,
                             padding=padding, output_padding=output_padding)
        self.num_experts, self.num_covars = num_experts, num_covars
        self.film = None
        adn = getattr(self, "adn", None)
        if num_experts == 1 and adn is not None and hasattr(adn, "N") and num_covars > 0:
            self.film = nn.Sequential(nn.Linear(num_covars, FILM_HIDDEN), nn.ReLU(),
                                      nn.Linear(FILM_HIDDEN, 2 * out_channels))
            nn.init.zeros_(self.film[2].weight)
            nn.init.zeros_(self.film[2].bias)

    def forward(self, x, covariate=None, out=None, defer=False):
        c = covariate_matrix(covariate, x) if covariate is not None and self.num_covars > 0 else None
        if self.num_experts > 1:
            r = torch.sigmoid(self.routing(c))                                       # [B, E]
            w = torch.einsum("be,eoi->boi", r, self.conv.weight.flatten(2))          # [B, Cout, Cin]
            b = r @ self.conv.bias if self.conv.bias is not None else None           # [B, Cout]
            if isinstance(x, ops.Deferred) and torch.is_grad_enabled() and (w.requires_grad or (b is not None and b.requires_grad)):
                x = x.materialize()
            y = ops.PerSampleConv1x1Fn.apply(x.raw, w, b, x) if isinstance(x, ops.Deferred) else ops.PerSampleConv1x1Fn.apply(x, w, b)
            if out is not None:
                ops._copy_channels(y, out)
                return out
            return y
        film = None
        if self.film is not None and c is not None:
            film = FilmBatch.active.get(id(self))      # filled by the model for this forward (no-grad: batched GEMMs, grad: fused MLP kernel)
            if film is None or film[0].shape[0] != c.shape[0]:
                film = self.film(c).chunk(2, dim=-1)
        return super().forward(x, film=film, out=out, defer=defer)


class CondConvBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size=3, strides=1, dropout=0.0, num_covars=0):
        super().__init__()
        common = dict(kernel_size=kernel_size, padding=None, adn_ordering="NDA", act="relu", norm=Norm.BATCH,
                      dropout=dropout, num_covars=num_covars)
        self.conv = nn.ModuleList([CondConvolution(spatial_dims, in_channels, out_channels, strides=strides, **common),
                                   CondConvolution(spatial_dims, out_channels, out_channels, strides=1, **common)])

    def forward(self, x, covariate=None):
        for layer in self.conv:
            x = layer(x, covariate)
        return x
