"""Host-side mirror of the MONAI blocks the reference builds on, executing through the CUDA kernels.

Module / parameter names follow MONAI (``conv``, ``adn.N/D/A``, ``W_g``/``W_x``/``psi``, ``up``,
``attention``/``upconv``/``merge``/``submodule``) so ``state_dict`` keys are those of checkpoints
written by the reference (attn_unet_data_parallel.py:946-953).  The child ``nn.Conv3d`` /
``nn.BatchNorm3d`` / ... objects are parameter containers only: forward never calls them, it hands
their tensors to ``ops`` (C ABI).  Tensors between blocks are NDHWC ``[B,D,H,W,C]``.

Reference interfaces mirrored: MONAI ``Convolution``/``ADN`` (monai/networks/blocks),
``attentionunet.ConvBlock/UpConv/AttentionBlock/AttentionLayer`` as used at
attn_unet_data_parallel.py:120-240,285-306,442,495-497,546-558.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class Norm:
    BATCH = "BATCH"
    INSTANCE = "INSTANCE"


# number of times each BatchNorm running-stat update is applied per forward (2 inside the backbone of
# ContrastiveAttentionUNET_DP: the reference runs it twice, attn_unet_data_parallel.py:664,666)
_bn_updates = 1


@contextlib.contextmanager
def bn_updates(n: int):
    global _bn_updates
    prev, _bn_updates = _bn_updates, n
    try:
        yield
    finally:
        _bn_updates = prev


def make_norm(spec, channels):
    name = str(spec[0] if isinstance(spec, (tuple, list)) else spec).upper()
    if name == "BATCH":
        return nn.BatchNorm3d(channels)
    if name == "INSTANCE":
        return nn.InstanceNorm3d(channels)
    raise ValueError(f"norm {spec!r} is not on the hot path")


def make_act(spec):
    kwargs = {}
    if isinstance(spec, (tuple, list)):
        spec, kwargs = spec[0], dict(spec[1])
    if isinstance(spec, type):
        return spec(**kwargs)
    name = str(spec).upper()
    if name == "PRELU":
        return nn.PReLU(**kwargs)
    if name == "RELU":
        return nn.ReLU(**kwargs)
    if name == "LEAKYRELU":
        return nn.LeakyReLU(**kwargs)
    raise ValueError(f"activation {spec!r} is not on the hot path")


def act_code(act: Optional[nn.Module]):
    """(COMA_ACT_*, slope tensor or None)"""
    if act is None:
        return L.ACT_NONE, None
    if isinstance(act, nn.ReLU):
        return L.ACT_RELU, None
    if isinstance(act, nn.PReLU):
        assert act.weight.numel() == 1
        return L.ACT_LEAKY, act.weight
    if isinstance(act, nn.LeakyReLU):
        return L.ACT_LEAKY, torch.full((1,), act.negative_slope, dtype=torch.float32)
    if isinstance(act, nn.Sigmoid):
        return L.ACT_SIGMOID, None
    raise ValueError(f"unsupported activation {act}")


class ADN(nn.Sequential):
    def __init__(self, ordering="NDA", in_channels=None, act="RELU", norm=None, dropout=None):
        super().__init__()
        parts = {"A": None, "D": None, "N": None}
        if norm is not None:
            parts["N"] = make_norm(norm, in_channels)
        if act is not None:
            parts["A"] = make_act(act)
        if dropout is not None:
            parts["D"] = nn.Dropout(p=float(dropout))
        for key in ordering.upper():
            if parts[key] is not None:
                self.add_module(key, parts[key])


def _grad_mode(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class Convolution(nn.Module):
    """MONAI ``Convolution``: conv (+ norm, dropout, activation in "NDA" order), on NDHWC tensors.

    ``film`` is an optional ``(dgamma, beta)`` pair of ``[B, Cout]`` tensors (CondConv modulation):
    y = act((1 + dgamma) * norm(conv(x)) + beta).  ``pad_out`` keeps zero channels in the stored
    output so the next conv sees a channel count the tcgen05 path accepts.
    """

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, adn_ordering="NDA",
                 act="PRELU", norm="INSTANCE", dropout=None, dropout_dim=1, dilation=1, groups=1, bias=True,
                 conv_only=False, is_transposed=False, padding=None, output_padding=None, pad_out=1):
        super().__init__()
        if spatial_dims != 3 or dilation != 1 or groups != 1:
            raise NotImplementedError("only 3-D, dilation 1, groups 1 convolutions are on the hot path")
        if kernel_size not in (1, 3) or strides not in (1, 2):
            raise NotImplementedError("kernel sizes 1/3 and strides 1/2 are on the hot path")
        if padding is not None and padding != (kernel_size - 1) // 2:
            raise NotImplementedError("'same' padding only")
        if adn_ordering.upper() != "NDA":
            raise NotImplementedError("adn_ordering other than NDA is not on the hot path")
        self.in_channels, self.out_channels, self.is_transposed = in_channels, out_channels, is_transposed
        self.kernel_size, self.strides, self.pad_out = kernel_size, strides, pad_out
        pad = (kernel_size - 1) // 2
        if is_transposed:
            if output_padding is not None and output_padding != strides - 1:
                raise NotImplementedError("output_padding must be stride-1")
            self.conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=strides, padding=pad,
                                           output_padding=strides - 1, bias=bias)
        else:
            self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=strides, padding=pad, bias=bias)
        if not conv_only and not (act is None and norm is None and dropout is None):
            self.adn = ADN(adn_ordering, out_channels, act, norm, dropout)
        self._slope_buf = None

    # -- helpers ---------------------------------------------------------------------------------
    def _parts(self):
        adn = getattr(self, "adn", None)
        norm = getattr(adn, "N", None) if adn is not None else None
        act = getattr(adn, "A", None) if adn is not None else None
        drop = getattr(adn, "D", None) if adn is not None else None
        if drop is not None and drop.p > 0 and self.training:
            raise NotImplementedError("dropout > 0 is not used by the reference configuration")
        return norm, act

    def _slope(self, act, device):
        code, slope = act_code(act)
        if slope is not None and not isinstance(slope, nn.Parameter):
            if self._slope_buf is None or self._slope_buf.device != device:
                self._slope_buf = slope.to(device)
            slope = self._slope_buf
        return code, slope

    def forward(self, x, film=None, out=None, final_relu=False, defer=False):
        """``x`` may be an ``ops.Deferred`` (its producer's norm/activation is applied by this conv's input prologue);
        ``defer=True`` (no-grad InstanceNorm path only) returns this conv's own output as a Deferred."""
        norm, act = self._parts()
        code, slope = self._slope(act, x.device)
        if final_relu:   # the model's final ReLU folded onto a PReLU head (attn_unet_data_parallel.py:654-656)
            assert code == L.ACT_LEAKY
            code = L.ACT_LEAKY_RELU
        conv = self.conv
        store = ops._round_up(self.out_channels, self.pad_out)
        B = x.shape[0]
        Cn = store
        # --- g, h of y = act(g * xhat + h) ------------------------------------------------------
        g = h = None
        mode = L.NORM_NONE
        bn_train = False
        if isinstance(norm, nn.modules.batchnorm._BatchNorm):
            bn_train = self.training or not norm.track_running_stats
            mode = L.NORM_BATCH if bn_train else L.NORM_GIVEN
            g, h = norm.weight, norm.bias
        elif isinstance(norm, nn.modules.instancenorm._InstanceNorm):
            mode = L.NORM_INSTANCE
            if norm.affine:
                g, h = norm.weight, norm.bias
        elif norm is not None:
            raise NotImplementedError(type(norm))
        if film is not None and norm is not None:
            dgamma, beta = film
            if g is None:
                g, h = 1 + dgamma, beta
            else:
                g, h = g[None, :] * (1 + dgamma), h[None, :] * (1 + dgamma) + beta
        if g is not None and Cn != self.out_channels:
            raise NotImplementedError("padded outputs with an affine norm")
        eps = norm.eps if norm is not None else 1e-5
        grad = _grad_mode(x, conv.weight, g if torch.is_tensor(g) else None)
        if grad:
            x = ops.materialized(x)

        # --- eval BatchNorm / no norm, no autograd: everything in the conv epilogue -----------------
        if not grad and mode in (L.NORM_NONE, L.NORM_GIVEN):
            if mode == L.NORM_NONE and code == L.ACT_NONE:
                scale = shift = None
            else:
                cfg = ops.NormCfg(mode=mode, act=code, eps=eps,
                                  running_mean=getattr(norm, "running_mean", None),
                                  running_var=getattr(norm, "running_var", None))
                dummy = (x.raw if isinstance(x, ops.Deferred) else x).new_empty(B, 1, 1, 1, Cn)
                scale, shift, _, _, _ = ops.norm_coefficients(dummy, g, h, cfg)
            wp, cout_comp = self._packed(x)
            if scale is not None and cout_comp != Cn:
                scale, shift = _pad_cols(scale, cout_comp, 1.0), _pad_cols(shift, cout_comp, 0.0)
            y, _ = ops.conv_raw(x, wp, conv.bias, ksize=self.kernel_size, stride=self.strides,
                                transposed=self.is_transposed, cout_store=store, scale=scale, shift=shift,
                                slope=None if slope is None else slope.detach(), act=code, out=out,
                                alg=(self.in_channels, self.out_channels))
            return y
        # --- general path: conv (+ statistics) then normalise / modulate / activate ------------------
        want_stats = mode in (L.NORM_INSTANCE, L.NORM_BATCH)
        pro = x if isinstance(x, ops.Deferred) else None
        cfg = ops.ConvCfg(ksize=self.kernel_size, stride=self.strides, transposed=self.is_transposed,
                          cout_store=store, want_stats=want_stats, bias_grad_zero=want_stats, in_affine=pro)
        y, stats = ops.conv3d(pro.raw if pro is not None else x, conv.weight, conv.bias, cfg)
        if norm is None and code == L.ACT_NONE:
            if out is not None:
                ops._copy_channels(y, out)
                return out
            return y
        ncfg = ops.NormCfg(mode=mode, act=code, eps=eps, stats=stats, out=out)
        if defer and not grad and out is None and mode == L.NORM_INSTANCE and code in (L.ACT_NONE, L.ACT_RELU, L.ACT_LEAKY):
            A, S, _, _, _ = ops.norm_coefficients(y, g, h, ncfg)
            sl = None if slope is None else slope.detach().float().reshape(-1)[:1].contiguous()
            return ops.Deferred(y, A, S, code, sl)
        if isinstance(norm, nn.modules.batchnorm._BatchNorm):
            ncfg.running_mean, ncfg.running_var = norm.running_mean, norm.running_var
            ncfg.momentum = 0.1 if norm.momentum is None else norm.momentum
            ncfg.update_running = bn_train and norm.track_running_stats and self.training
            ncfg.n_updates = _bn_updates
            if ncfg.update_running:
                norm.num_batches_tracked += _bn_updates
        return ops.norm_act(y, g, h, slope, ncfg)

    def _packed(self, x):
        cin_buf = x.shape[-1]
        w = self.conv.weight
        cout_w = self.out_channels
        store = ops._round_up(cout_w, self.pad_out)
        use_tc_pad = x.dtype == torch.bfloat16 and cin_buf % 16 == 0
        cout_comp = max(ops._round_up(cout_w, 16) if use_tc_pad else cout_w, store)
        return ops.pack_weight(w, self.is_transposed, cin_buf, cout_comp, x.dtype), cout_comp


def _pad_cols(t, n, fill):
    out = t.new_full((t.shape[0], n), fill)
    out[:, :t.shape[1]] = t
    return out


class ConvBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size=3, strides=1, dropout=0.0):
        super().__init__()
        common = dict(kernel_size=kernel_size, padding=None, adn_ordering="NDA", act="relu", norm=Norm.BATCH,
                      dropout=dropout)
        self.conv = nn.Sequential(Convolution(spatial_dims, in_channels, out_channels, strides=strides, **common),
                                  Convolution(spatial_dims, out_channels, out_channels, strides=1, **common))

    def forward(self, x):
        return self.conv(x)


class UpConv(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size=3, strides=2, dropout=0.0):
        super().__init__()
        self.up = Convolution(spatial_dims, in_channels, out_channels, strides=strides, kernel_size=kernel_size,
                              act="relu", adn_ordering="NDA", norm=Norm.BATCH, dropout=dropout, is_transposed=True)

    def forward(self, x, out=None):
        return self.up(x, out=out)


class AttentionBlock(nn.Module):
    """out = x * sigmoid(BN(psi(relu(BN(W_g g) + BN(W_x x)))))."""

    FUSED_C = (8, 16, 32, 64)

    def __init__(self, spatial_dims, f_int, f_g, f_l, dropout=0.0):
        super().__init__()

        def pointwise(cin, cout):
            return Convolution(spatial_dims, cin, cout, kernel_size=1, strides=1, padding=0, dropout=dropout,
                               conv_only=True)

        self.W_g = nn.Sequential(pointwise(f_g, f_int), nn.BatchNorm3d(f_int))
        self.W_x = nn.Sequential(pointwise(f_l, f_int), nn.BatchNorm3d(f_int))
        self.psi = nn.Sequential(pointwise(f_int, 1), nn.BatchNorm3d(1), nn.Sigmoid())
        self.relu = nn.ReLU()
        self._fold_key, self._folded = None, None

    # BatchNorm (running statistics) folded into the three 1x1x1 convolutions
    def _fold(self):
        mods = (self.W_g[0].conv, self.W_g[1], self.W_x[0].conv, self.W_x[1], self.psi[0].conv, self.psi[1])
        key = tuple(t._version for m in mods for t in list(m.parameters()) + list(m.buffers())) + \
            (self.W_g[0].conv.weight.device, ops.weight_epoch())
        if key == self._fold_key:
            return self._folded
        with torch.no_grad():
            def fold(conv, bn):
                s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                w = conv.weight.reshape(conv.weight.shape[0], -1) * s[:, None]
                b = (conv.bias - bn.running_mean) * s + bn.bias
                return w.float().contiguous(), b.float().contiguous()
            wg, bg = fold(self.W_g[0].conv, self.W_g[1])
            wx, bx = fold(self.W_x[0].conv, self.W_x[1])
            wp, bp = fold(self.psi[0].conv, self.psi[1])
            self._folded = (wg, wx, (bg + bx).contiguous(), wp.reshape(-1).contiguous(), bp.reshape(1).contiguous())
        self._fold_key = key
        return self._folded

    def _bn(self, t_raw, stats, bn, act, residual=None, out=None):
        train = self.training
        cfg = ops.NormCfg(mode=L.NORM_BATCH if train else L.NORM_GIVEN, act=act, eps=bn.eps, stats=stats,
                          running_mean=bn.running_mean, running_var=bn.running_var,
                          momentum=0.1 if bn.momentum is None else bn.momentum, update_running=train,
                          n_updates=_bn_updates, out=out)
        if train:
            bn.num_batches_tracked += _bn_updates
        return ops.norm_act(t_raw, bn.weight, bn.bias, None, cfg, residual=residual)

    def forward(self, g, x, out=None, want_coeff=False):
        Cn = x.shape[-1]
        grad = _grad_mode(g, x, self.W_g[0].conv.weight)
        if not self.training and not grad and Cn in self.FUSED_C:
            wg, wx, bsum, wpsi, bpsi = self._fold()
            psi = x.new_empty(*x.shape[:-1], 1) if want_coeff else None
            y = ops.gate_fused(g, x, wg, wx, bsum, wpsi, bpsi, out=out, psi_out=psi)
            return (y, psi) if want_coeff else y
        # composed path (training statistics, or channel counts the fused kernel does not take)
        st = self.training
        cg, cx, cp = self.W_g[0].conv, self.W_x[0].conv, self.psi[0].conv
        k1 = dict(ksize=1, stride=1, want_stats=st, bias_grad_zero=st)
        tg, sg = ops.conv3d(g, cg.weight, cg.bias, ops.ConvCfg(**k1))
        tx, sx = ops.conv3d(x, cx.weight, cx.bias, ops.ConvCfg(**k1))
        u = self._bn(tg, sg, self.W_g[1], L.ACT_NONE)
        s = self._bn(tx, sx, self.W_x[1], L.ACT_RELU, residual=u)
        q, sq = ops.conv3d(s, cp.weight, cp.bias, ops.ConvCfg(**k1))
        p = self._bn(q, sq, self.psi[1], L.ACT_SIGMOID)
        y = ops.bcast_mul(x, p, out)
        return (y, p) if want_coeff else y


class AttentionLayer(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, submodule, up_kernel_size=3, strides=2, dropout=0.0):
        super().__init__()
        self.attention = AttentionBlock(spatial_dims, f_g=in_channels, f_l=in_channels, f_int=in_channels // 2)
        self.upconv = UpConv(spatial_dims, out_channels, in_channels, strides=strides, kernel_size=up_kernel_size)
        self.merge = Convolution(spatial_dims, 2 * in_channels, in_channels, dropout=dropout)
        self.submodule = submodule
