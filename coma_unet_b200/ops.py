"""Autograd-aware Python wrappers over the C ABI (include/coma_b200.h).

Activations are torch tensors shaped ``[B, D, H, W, C]`` (NDHWC); a tensor may be a channel slice of
a wider buffer (its voxel stride is then the buffer's channel count), which is how concat buffers
are written in place.  Parameters stay fp32 in the reference's layouts; weights are re-packed to
``[tap][Cout][Cin]`` in the activation dtype (cached per parameter version).

Every volume-sized operation is a CUDA kernel of this repo; only O(B*C)-sized vector math (FiLM MLP
outputs, BatchNorm folding) is done with torch ops.
"""
from __future__ import annotations

import contextlib
import os

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib as L


# ------------------------------------------------------------------------------------------------
# NDHWC views
# ------------------------------------------------------------------------------------------------
def vol_cs(t: torch.Tensor) -> int:
    """Voxel stride (elements) of an NDHWC tensor that may be a channel slice of a wider buffer."""
    assert t.dim() == 5, f"expected [B,D,H,W,C], got {tuple(t.shape)}"
    if t.is_contiguous():
        return t.shape[4]
    B, D, H, W, Cn = t.shape
    if Cn > 1:
        assert t.stride(4) == 1, "channels must be innermost"
    cs = None
    for dim, inner in ((3, 1), (2, W), (1, W * H), (0, W * H * D)):
        if t.shape[dim] > 1:
            cs = t.stride(dim) // inner
            break
    if cs is None:
        cs = Cn
    for dim, inner in ((3, 1), (2, W), (1, W * H), (0, W * H * D)):
        if t.shape[dim] > 1:
            assert t.stride(dim) == cs * inner, f"not an NDHWC (channel-sliced) layout: {t.shape} {t.stride()}"
    assert cs >= Cn
    return cs


def as_vol(t: torch.Tensor) -> torch.Tensor:
    """Make sure ``t`` is a valid NDHWC tensor (copies only when the strides are not usable)."""
    try:
        vol_cs(t)
        return t
    except AssertionError:
        return t.contiguous()


def ncdhw_to_vol(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """User-facing ``[B,C,D,H,W]`` -> internal ``[B,D,H,W,C]`` in ``dtype`` (free for C == 1)."""
    B, Cn, D, H, W = x.shape
    if Cn == 1:
        return x.reshape(B, D, H, W, 1).to(dtype)
    return x.permute(0, 2, 3, 4, 1).to(dtype).contiguous()


def vol_to_ncdhw(t: torch.Tensor) -> torch.Tensor:
    """Internal NDHWC -> logical ``[B,C,D,H,W]`` view (channels_last_3d strides, no copy)."""
    return t.permute(0, 4, 1, 2, 3)


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


# ------------------------------------------------------------------------------------------------
# weight packing (cached)
# ------------------------------------------------------------------------------------------------
# Derived-weight caches (packed conv weights, folded BatchNorm, stacked FiLM MLPs) are keyed by the parameters' version counters
# AND by this epoch.  The version counter alone is not enough: torch's FUSED optimizers (``AdamW(fused=True)``, which train_dp
# uses) update parameters in place without bumping it -- measured: with the version-only key the convolutions kept using the
# weights packed in the first step while every directly-read parameter trained.  The model classes bump the epoch at the start
# of every training-mode forward and on every train()/eval() switch, so a pack is reused only within one training step
# (forward + backward) or across consecutive eval-mode forwards.
_EPOCH = 0


def bump_weight_epoch() -> None:
    global _EPOCH
    _EPOCH += 1


def weight_epoch() -> int:
    return _EPOCH


def pack_weight(weight: torch.Tensor, transposed: bool, cin_buf: int, cout_comp: int, dtype: torch.dtype) -> torch.Tensor:
    """Conv3d ``[Cout,Cin,k,k,k]`` / ConvTranspose3d ``[Cin,Cout,k,k,k]`` -> ``[k^3, cout_comp, cin_buf]`` (zero padded).

    Cached on the parameter object itself (so the cache dies with it) and keyed by its version counter, its storage and the
    weight epoch (see above)."""
    cache = weight.__dict__.setdefault("_coma_packed", {})
    key = (transposed, cin_buf, cout_comp, dtype)
    hit = cache.get(key)
    if hit is not None and hit[0] == (weight._version, weight.data_ptr(), weight.device, _EPOCH):
        return hit[1]
    w = weight.detach()
    k = w.shape[2]
    if _layout_kernel_ok(w, dtype):
        p = torch.empty(k ** 3, cout_comp, cin_buf, device=w.device, dtype=dtype)
        _weight_layout(w, p, swap=transposed)
        cache[key] = ((weight._version, weight.data_ptr(), weight.device, _EPOCH), p)
        return p
    if transposed:
        p = w.permute(2, 3, 4, 1, 0)
    else:
        p = w.permute(2, 3, 4, 0, 1)
    p = p.reshape(k ** 3, p.shape[3], p.shape[4])
    if p.shape[1] != cout_comp or p.shape[2] != cin_buf:
        q = torch.zeros(k ** 3, cout_comp, cin_buf, device=w.device, dtype=w.dtype)
        q[:, :p.shape[1], :p.shape[2]] = p
        p = q
    p = p.to(dtype).contiguous()
    cache[key] = ((weight._version, weight.data_ptr(), weight.device, _EPOCH), p)
    return p


PACK_KERNEL = os.environ.get("COMA_DISABLE_PACK_KERNEL", "0") != "1"


def _layout_kernel_ok(w: torch.Tensor, dtype: torch.dtype) -> bool:
    return (PACK_KERNEL and w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 5
            and w.shape[2] == w.shape[3] == w.shape[4] and w.shape[2] ** 3 <= 27 and dtype in (torch.float32, torch.bfloat16))


def _weight_layout(param, packed, swap=False, flip=False, unpack=False):
    """coma_weight_layout: parameter tensor [A][B][k,k,k] fp32 <-> tap-major [T][rows][cols] (one launch)."""
    a = L.WeightLayoutArgs()
    a.param, a.packed = L.ptr(param), L.ptr(packed)
    a.A, a.B, a.T = param.shape[0], param.shape[1], packed.shape[0]
    a.R_pad, a.C_pad = packed.shape[1], packed.shape[2]
    a.swap, a.flip, a.dtype, a.unpack = int(swap), int(flip), L.dtype_code(packed.dtype), int(unpack)
    L.call("coma_weight_layout", C.byref(a), L.stream())


def pack_adjoint(weight: torch.Tensor, transposed: bool, stride: int, cin_buf: int, cout_comp: int, cout_store: int,
                 dtype: torch.dtype) -> torch.Tensor:
    """The data-gradient operand ``[k^3, cin_buf, cout_store]`` of a conv whose forward operand is ``pack_weight(...)``: channels
    swapped, taps flipped for a stride-1 convolution (its adjoint is a convolution with the mirrored kernel; the adjoint of a
    strided conv is a transposed conv over the same taps and vice versa).  Cached like pack_weight."""
    flip = not transposed and stride == 1
    if not _layout_kernel_ok(weight, dtype):
        wp = pack_weight(weight, transposed, cin_buf, cout_comp, dtype)[:, :cout_store, :]
        return (wp.flip(0) if flip else wp).transpose(1, 2).contiguous()
    cache = weight.__dict__.setdefault("_coma_packed", {})
    key = ("adjoint", transposed, stride, cin_buf, cout_store, dtype)
    stamp = (weight._version, weight.data_ptr(), weight.device, _EPOCH)
    hit = cache.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    w = weight.detach()
    p = torch.empty(w.shape[2] ** 3, cin_buf, cout_store, device=w.device, dtype=dtype)
    _weight_layout(w, p, swap=not transposed, flip=flip)
    cache[key] = (stamp, p)
    return p


def unpack_weight_gradient(dwp: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """``[k^3, rows, cols]`` as the weight-gradient kernels write it -> the parameter's own layout ``[rows_w, cols_w, k, k, k]``."""
    A, B, k = like.shape[0], like.shape[1], like.shape[2]
    if _layout_kernel_ok(like, torch.float32) and dwp.dtype == torch.float32 and dwp.is_contiguous():
        dw = torch.empty(like.shape, device=dwp.device, dtype=torch.float32)
        _weight_layout(dw, dwp, unpack=True)
        return dw
    return dwp[:, :A, :B].reshape(k, k, k, A, B).permute(3, 4, 0, 1, 2).contiguous()


def invalidate_weight_caches(module: torch.nn.Module) -> None:
    """Drop every cached packed weight / folded BatchNorm of ``module``.

    The caches are keyed by the parameters' version counters and the weight epoch (``bump_weight_epoch``).  In-place updates
    through ``.data`` (EMA swaps, manual clipping on ``p.data``) and fused optimizers bump neither: after such an update of a
    model that stays in eval mode, call this (training-mode forwards and train()/eval() switches refresh the caches by
    themselves)."""
    bump_weight_epoch()
    for p in module.parameters():
        p.__dict__.pop("_coma_packed", None)
    for m in module.modules():
        if hasattr(m, "_fold_key"):
            m._fold_key = None
        if hasattr(m, "_fb"):
            m._fb = None


def _out_extent(n: int, k: int, stride: int, transposed: bool) -> int:
    if transposed:
        return n * stride
    pad = (k - 1) // 2
    return (n + 2 * pad - k) // stride + 1


def _conv_args(x, wp, bias, y, *, ksize, stride, transposed, cout_comp, scale=None, shift=None, slope=None, stats=None,
               act=L.ACT_NONE, impl=L.IMPL_AUTO, w_bstride=0, bias_bstride=0, alg=None) -> L.ConvArgs:
    B, Di, Hi, Wi, Cin = x.shape
    _, Do, Ho, Wo, ycn = y.shape
    a = L.ConvArgs()
    a.x, a.w, a.bias, a.y = L.ptr(x), L.ptr(wp), L.ptr(bias), L.ptr(y)
    a.scale, a.shift, a.slope, a.stats = L.ptr(scale), L.ptr(shift), L.ptr(slope), L.ptr(stats)
    a.B, a.Di, a.Hi, a.Wi, a.Do, a.Ho, a.Wo = B, Di, Hi, Wi, Do, Ho, Wo
    a.Cin, a.Cout = Cin, cout_comp
    a.x_cs, a.x_co, a.y_cs, a.y_co, a.y_cn = vol_cs(x), 0, vol_cs(y), 0, ycn
    a.ksize, a.stride, a.pad, a.transposed = ksize, stride, (ksize - 1) // 2, int(transposed)
    a.w_bstride, a.bias_bstride = w_bstride, bias_bstride
    a.act, a.dtype, a.impl = act, L.dtype_code(x.dtype), impl
    # (Cin, Cout) of the LAYER, without the zero padding the tensor-core kernels compute on: bench.py counts algorithmic FLOPs
    # from these (a Python attribute of the ctypes object, not part of the C struct)
    a.alg = alg if alg is not None else (Cin, cout_comp)
    return a


def _run_conv(a: L.ConvArgs, kind: str):
    L.call(kind, C.byref(a), L.stream())


@dataclass
class Deferred:
    """A conv output whose norm + FiLM + activation has not been applied yet: ``act(A[b,c] * raw + S[b,c])``.
    A consumer conv that supports the input prologue (include/coma_b200.h, coma_conv_args.in_scale) applies it on load;
    anything else calls ``materialize()`` (one coma_norm_film_act_fwd sweep, what the producer would have run)."""
    raw: torch.Tensor
    A: torch.Tensor
    S: torch.Tensor
    act: int
    slope: Optional[torch.Tensor]

    shape = property(lambda self: self.raw.shape)
    dtype = property(lambda self: self.raw.dtype)
    device = property(lambda self: self.raw.device)
    requires_grad = False

    def materialize(self, out=None):
        return affine_act(self.raw, self.A, self.S, self.slope, self.act, out=out)


def taps_conv_ok(x) -> bool:
    """The tap-packed few-channel conv kernel needs planes of at least 16 x 8 voxels (bf16)."""
    return x.dtype == torch.bfloat16 and x.shape[2] >= 16 and x.shape[3] >= 8


def materialized(x):
    return x.materialize() if isinstance(x, Deferred) else x


# ------------------------------------------------------------------------------------------------
# fp32 tensors on the tensor cores: split-precision 3x3x3 convolutions
# ------------------------------------------------------------------------------------------------
# With ``compute_dtype=torch.float32`` activations and packed weights are fp32 and the convolutions run on CUDA cores (exact).
# ``fp32_split(True)`` (model kwarg ``fp32_tensor_cores=True``) moves the 3x3x3 convolutions -- 97 % of the FLOPs -- onto tcgen05 at
# fp32-level accuracy: x = x_hi + x_lo, w = w_hi + w_lo with bf16 halves (2^-17 relative residual per operand),
#     conv(x, w) ~= conv([x_hi | x_lo | x_hi], [w_hi | w_hi | w_lo])          (the lo x lo term, 2^-18, is dropped)
# i.e. ONE bf16 convolution over three times the input channels whose fp32 accumulator is stored unrounded
# (dtype COMA_BF16_F32OUT, include/coma_b200.h); weight gradients likewise on [g_hi | g_lo] x [x_hi | x_lo] (three of the four
# blocks of the result are summed).  Everything else of the fp32 path is unchanged.
FP32_SPLIT = False


@contextlib.contextmanager
def fp32_split(on: bool = True):
    global FP32_SPLIT
    prev, FP32_SPLIT = FP32_SPLIT, bool(on)
    try:
        yield
    finally:
        FP32_SPLIT = prev


def _hi_lo(t: torch.Tensor):
    hi = t.to(torch.bfloat16)
    return hi, (t - hi.float()).to(torch.bfloat16)


def _cat_padded(parts, width):
    """Channel-concatenate bf16 tensors, each zero-padded to ``width`` channels."""
    n = parts[0].shape[-1]
    if n == width:
        return torch.cat(parts, dim=-1)
    out = parts[0].new_zeros(*parts[0].shape[:-1], width * len(parts))
    for i, p in enumerate(parts):
        out[..., i * width:i * width + n] = p
    return out


def _conv_raw_split(x, wp, bias, *, ksize, stride, transposed, cout_store, scale, shift, slope, act, want_stats, out, kind, alg):
    B, Di, Hi, Wi, Cin = x.shape
    taps, cout_comp, _ = wp.shape
    cin_p, cout_p = _round_up(Cin, 16), _round_up(cout_comp, 16)
    xh, xl = _hi_lo(x)
    xs = _cat_padded([xh, xl, xh], cin_p)
    wh, wl = _hi_lo(wp)
    ws = torch.zeros(taps, cout_p, 3 * cin_p, device=x.device, dtype=torch.bfloat16)
    ws[:, :cout_comp, :Cin], ws[:, :cout_comp, cin_p:cin_p + Cin], ws[:, :cout_comp, 2 * cin_p:2 * cin_p + Cin] = wh, wh, wl
    cout_store = cout_store or cout_comp
    Do, Ho, Wo = (_out_extent(n, ksize, stride, transposed) for n in (Di, Hi, Wi))
    y = out if out is not None else torch.empty(B, Do, Ho, Wo, cout_store, device=x.device, dtype=torch.float32)

    def padded(t, fill):
        if t is None or t.shape[-1] == cout_p:
            return None if t is None else t.float().contiguous()
        q = torch.full((*t.shape[:-1], cout_p), fill, device=x.device, dtype=torch.float32)
        q[..., :t.shape[-1]] = t
        return q

    a = _conv_args(xs, ws, padded(bias, 0.0), y, ksize=ksize, stride=stride, transposed=transposed, cout_comp=cout_p,
                   scale=padded(scale, 1.0), shift=padded(shift, 0.0), slope=slope, act=act, alg=alg or (Cin, cout_comp))
    a.dtype = L.BF16_F32OUT
    stats = None
    if want_stats:
        chunks = L.lib().coma_conv3d_stat_chunks(C.byref(a))
        stats = torch.empty(B, chunks, cout_p, 2, device=x.device, dtype=torch.float32)
        a.stats = L.ptr(stats)
    _run_conv(a, kind or ("coma_convT3d_fprop" if transposed else "coma_conv3d_fprop"))
    if stats is not None and cout_p != cout_comp:
        stats = stats[:, :, :cout_comp, :].contiguous()
    return y, stats


def conv_raw(x, wp, bias, *, ksize, stride=1, transposed=False, cout_store=None, scale=None, shift=None, slope=None,
             act=L.ACT_NONE, want_stats=False, out=None, impl=L.IMPL_AUTO, w_bstride=0, bias_bstride=0, kind=None, alg=None):
    """One conv kernel launch on packed weights.  Returns (y, stats_partial or None).  ``x`` may be a Deferred."""
    pro = x if isinstance(x, Deferred) else None
    x = as_vol(pro.raw if pro is not None else x)
    if (FP32_SPLIT and x.dtype == torch.float32 and pro is None and ksize == 3 and not w_bstride and not bias_bstride
            and impl == L.IMPL_AUTO and act != L.ACT_SIGMOID):
        return _conv_raw_split(x, wp, bias, ksize=ksize, stride=stride, transposed=transposed, cout_store=cout_store, scale=scale,
                               shift=shift, slope=slope, act=act, want_stats=want_stats, out=out, kind=kind, alg=alg)
    B, Di, Hi, Wi, Cin = x.shape
    per = wp.shape[-3:] if w_bstride else wp.shape
    taps, cout_comp, cin_w = per
    assert cin_w == Cin and taps == ksize ** 3, (wp.shape, x.shape, ksize)
    cout_store = cout_store or cout_comp
    Do, Ho, Wo = (_out_extent(n, ksize, stride, transposed) for n in (Di, Hi, Wi))
    y = out if out is not None else torch.empty(B, Do, Ho, Wo, cout_store, device=x.device, dtype=x.dtype)
    assert tuple(y.shape) == (B, Do, Ho, Wo, cout_store)
    if bias is not None and bias.shape[-1] != cout_comp:
        pb = torch.zeros(*bias.shape[:-1], cout_comp, device=x.device, dtype=torch.float32)
        pb[..., :bias.shape[-1]] = bias
        bias = pb
    a = _conv_args(x, wp, None if bias is None else bias.float().contiguous(), y, ksize=ksize, stride=stride,
                   transposed=transposed, cout_comp=cout_comp, scale=scale, shift=shift, slope=slope, act=act, impl=impl,
                   w_bstride=w_bstride, bias_bstride=bias_bstride, alg=alg)
    if pro is not None:
        a.in_scale, a.in_shift, a.in_slope, a.in_act = L.ptr(pro.A), L.ptr(pro.S), L.ptr(pro.slope), pro.act
        if pro.act not in (L.ACT_NONE, L.ACT_RELU, L.ACT_LEAKY) or not L.lib().coma_conv3d_prologue_supported(C.byref(a)):
            x = pro.materialize()                      # no fused prologue for this shape: apply it the ordinary way
            a.x, a.x_cs = L.ptr(x), vol_cs(x)
            a.in_scale = a.in_shift = a.in_slope = None
            a.in_act = 0
    stats = None
    if want_stats:
        chunks = L.lib().coma_conv3d_stat_chunks(C.byref(a))
        stats = torch.empty(B, chunks, cout_comp, 2, device=x.device, dtype=torch.float32)
        a.stats = L.ptr(stats)
    _run_conv(a, kind or ("coma_convT3d_fprop" if transposed else "coma_conv3d_fprop"))
    return y, stats


DETERMINISTIC_WGRAD = os.environ.get("COMA_WGRAD_ATOMICS", "0") != "1"     # A/B switch: fp32 atomics into a zeroed dw instead
_CG1_PAD = os.environ.get("COMA_CG1_PAD", "0") == "1"       # A/B switch: one-channel k3 weight gradients zero-padded to the tcgen05 16-row instance
_CG1_SIMT = os.environ.get("COMA_CG1_SIMT", "0") == "1"     # A/B switch: keep one-channel weight gradients on the CUDA-core sweep


def wgrad_raw(g, x, *, ksize, stride, kind="coma_conv3d_wgrad", alg=None):
    """dw[tap][Cg][Cx] = sum_o g[o] (x) x[o*stride + k - pad]  (conv geometry, fp32)."""
    g, x = as_vol(g), as_vol(x)
    B, Dg, Hg, Wg, Cg = g.shape
    _, Dx, Hx, Wx, Cx = x.shape
    if FP32_SPLIT and g.dtype == torch.float32 and ksize == 3:
        # split-precision weight gradient: [g_hi | g_lo] x [x_hi | x_lo] in bf16, the hi*hi + lo*hi + hi*lo blocks summed in fp32
        cgp, cxp = _round_up(Cg, 16), _round_up(Cx, 16)
        g2, x2 = _cat_padded(list(_hi_lo(g)), cgp), _cat_padded(list(_hi_lo(x)), cxp)
        d2 = wgrad_raw(g2, x2, ksize=ksize, stride=stride, kind=kind, alg=alg or (Cg, Cx))
        return (d2[:, :Cg, :Cx] + d2[:, cgp:cgp + Cg, :Cx] + d2[:, :Cg, cxp:cxp + Cx]).contiguous()
    if (Cg == 1 and ksize == 3 and stride == 1 and g.dtype == torch.bfloat16 and Cx % 16 == 0 and not _CG1_SIMT
            and not (Cx == 16 and Wg % 32 == 0 and Hg % 8 == 0 and not _CG1_PAD)       # the gathered-A mma.sync kernel takes these
            and ((Wg % 32 == 0 and Hg % 8 == 0) or (Wg == 16 and Hg % 16 == 0)) and Dg >= 4):
        # one-channel gradient (the 16 -> 1 modulator heads): the CUDA-core sweep is latency-bound (1.1 ms for 0.3 GB at batch 4);
        # zero-padded to one 16-channel row it runs on the tcgen05 weight-gradient kernel (0.1 ms pad + 0.3 ms)
        g16 = torch.nn.functional.pad(g, (0, 15))
        return wgrad_raw(g16, x, ksize=ksize, stride=stride, kind=kind, alg=alg or (1, Cx))[:, :1, :].contiguous()
    a = L.WgradArgs()
    a.g, a.x = L.ptr(g), L.ptr(x)
    a.B, a.Dg, a.Hg, a.Wg, a.Dx, a.Hx, a.Wx = B, Dg, Hg, Wg, Dx, Hx, Wx
    a.Cg, a.Cx, a.g_cs, a.g_co, a.x_cs, a.x_co = Cg, Cx, vol_cs(g), 0, vol_cs(x), 0
    a.ksize, a.stride, a.pad, a.dtype, a.impl = ksize, stride, (ksize - 1) // 2, L.dtype_code(g.dtype), L.IMPL_AUTO
    a.alg = alg if alg is not None else (Cg, Cx)       # layer channel counts without padding (see _conv_args)
    ws_bytes = L.lib().coma_conv3d_wgrad_workspace_size(C.byref(a)) if DETERMINISTIC_WGRAD else 0
    if ws_bytes > 0:
        # tcgen05 kernels, deterministic finish: per-CTA partial blocks summed in a fixed order and stored (no zero fill, no atomics)
        ws = torch.empty(ws_bytes, device=g.device, dtype=torch.uint8)
        dw = torch.empty(ksize ** 3, Cg, Cx, device=g.device, dtype=torch.float32)
        a.workspace, a.workspace_bytes = L.ptr(ws), ws_bytes
    else:
        dw = torch.zeros(ksize ** 3, Cg, Cx, device=g.device, dtype=torch.float32)
    a.dw = L.ptr(dw)
    L.call(kind, C.byref(a), L.stream())
    return dw


def channel_sums(t: torch.Tensor) -> torch.Tensor:
    """[B, C] sums over voxels (used for bias gradients); one coma_norm_stats sweep."""
    t = as_vol(t)
    B, D, H, W, Cn = t.shape
    V = D * H * W
    chunks = L.lib().coma_norm_stats_chunks(V)
    part = torch.empty(B, chunks, Cn, 2, device=t.device, dtype=torch.float32)
    L.call("coma_norm_stats", L.ptr(t), B, V, Cn, vol_cs(t), 0, L.dtype_code(t.dtype), L.ptr(part), L.stream())
    return part[..., 0].sum(dim=1)


@dataclass
class ConvCfg:
    ksize: int = 3
    stride: int = 1
    transposed: bool = False
    cout_store: Optional[int] = None   # channels of the output tensor (>= weight's Cout: zero padding)
    want_stats: bool = False
    bias_grad_zero: bool = False       # a mean-subtracting norm follows: d bias == 0 exactly
    impl: int = L.IMPL_AUTO
    in_affine: Optional["Deferred"] = None   # input prologue (no-grad paths)


class ConvFn(torch.autograd.Function):
    """Conv3d / ConvTranspose3d on fp32 master weights; backward = adjoint conv + weight-gradient kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, cfg: ConvCfg):
        pro = cfg.in_affine            # Deferred producer (no-grad only): x is its raw tensor
        x = as_vol(x)
        cin_buf = x.shape[-1]
        cout_w = weight.shape[1] if cfg.transposed else weight.shape[0]
        cout_store = cfg.cout_store or cout_w
        # one-output pointwise convs (the gate's psi, ProjectionHead 0) stream on the conv_pw1 kernel: padded to 16 outputs for the
        # tensor-core tile kernel they cost 0.85 ms instead of 0.1-0.17 ms at 128^3, batch 4
        pw1 = (cfg.ksize == 1 and cfg.stride == 1 and not cfg.transposed and cout_w == 1 and cout_store == 1
               and cin_buf in (1, 2, 3, 8, 16, 32, 64))
        use_tc_pad = x.dtype == torch.bfloat16 and cin_buf % 16 == 0 and cfg.impl != L.IMPL_SIMT and not pw1
        cout_comp = max(_round_up(cout_w, 16) if use_tc_pad else cout_w, cout_store)
        wp = pack_weight(weight, cfg.transposed, cin_buf, cout_comp, x.dtype)
        cin_w = weight.shape[0] if cfg.transposed else weight.shape[1]
        y, stats = conv_raw(pro if pro is not None else x, wp, bias, ksize=cfg.ksize, stride=cfg.stride, transposed=cfg.transposed,
                            cout_store=cout_store, want_stats=cfg.want_stats, impl=cfg.impl, alg=(cin_w, cout_w))
        ctx.save_for_backward(x, weight)
        ctx.cfg, ctx.cout_comp, ctx.has_bias = cfg, cout_comp, bias is not None
        ctx.split = FP32_SPLIT
        if stats is None:
            stats = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        with fp32_split(ctx.split):
            return ConvFn._backward(ctx, dy)

    @staticmethod
    def _backward(ctx, dy):
        x, weight = ctx.saved_tensors
        cfg: ConvCfg = ctx.cfg
        dy = as_vol(dy if dy.dtype == x.dtype else dy.to(x.dtype))
        cin_buf, cout_store = x.shape[-1], dy.shape[-1]
        k = cfg.ksize
        dx = dw = db = None
        cin_w, cout_w = (weight.shape[0], weight.shape[1]) if cfg.transposed else (weight.shape[1], weight.shape[0])
        if ctx.needs_input_grad[0]:
            adj = pack_adjoint(weight, cfg.transposed, cfg.stride, cin_buf, ctx.cout_comp, cout_store, x.dtype)
            alg = (cout_w, cin_w)
            if cfg.transposed:      # adjoint of convT(stride s) = conv(stride s), channels swapped
                dx, _ = conv_raw(dy, adj, None, ksize=k, stride=cfg.stride, transposed=False, kind="coma_convT3d_dgrad", alg=alg)
            elif cfg.stride == 1:   # adjoint of conv(stride 1) = conv with flipped taps, channels swapped
                dx, _ = conv_raw(dy, adj, None, ksize=k, stride=1, transposed=False, kind="coma_conv3d_dgrad", alg=alg)
            else:                   # adjoint of conv(stride s) = convT(stride s)
                dx, _ = conv_raw(dy, adj, None, ksize=k, stride=cfg.stride, transposed=True, kind="coma_conv3d_dgrad", alg=alg)
            assert dx.shape == x.shape, (dx.shape, x.shape)
        if ctx.needs_input_grad[1]:
            if cfg.transposed:
                dwp = wgrad_raw(x, dy, ksize=k, stride=cfg.stride, kind="coma_convT3d_wgrad", alg=(cin_w, cout_w))   # [T, cin_buf, cout_store]
                dw = unpack_weight_gradient(dwp, weight)
            else:
                dwp = wgrad_raw(dy, x, ksize=k, stride=cfg.stride, alg=(cout_w, cin_w))           # [T, cout_store, cin_buf]
                dw = unpack_weight_gradient(dwp, weight)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            if cfg.bias_grad_zero:
                db = torch.zeros(cout_w, device=x.device, dtype=torch.float32)
            else:
                db = channel_sums(dy).sum(dim=0)[:cout_w]
        return dx, dw, db, None


def conv3d(x, weight, bias, cfg: ConvCfg):
    y, stats = ConvFn.apply(x, weight, bias, cfg)
    return y, (stats if cfg.want_stats else None)


class PerSampleConv1x1Fn(torch.autograd.Function):
    """1x1x1 conv with per-sample weights ``w[B,Cout,Cin]`` / bias ``b[B,Cout]`` (expert-mixed reduce_channels)."""

    @staticmethod
    def forward(ctx, x, w, b, pro=None):
        x = as_vol(x)
        B, Cout, Cin = w.shape
        wp = w.detach().to(x.dtype).reshape(B, 1, Cout, Cin).contiguous()
        bb = None if b is None else b.detach().float().contiguous()
        y, _ = conv_raw(pro if pro is not None else x, wp, bb, ksize=1, w_bstride=Cout * Cin,
                        bias_bstride=Cout if b is not None else 0, impl=L.IMPL_SIMT)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = as_vol(dy if dy.dtype == x.dtype else dy.to(x.dtype))
        B, Cout, Cin = w.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            adj = w.detach().transpose(1, 2).to(x.dtype).reshape(B, 1, Cin, Cout).contiguous()
            dx, _ = conv_raw(dy, adj, None, ksize=1, w_bstride=Cout * Cin, impl=L.IMPL_SIMT, kind="coma_conv3d_dgrad")
        if ctx.needs_input_grad[1]:
            dw = torch.stack([wgrad_raw(dy[i:i + 1], x[i:i + 1], ksize=1, stride=1)[0] for i in range(B)])
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = channel_sums(dy)
        return dx, dw, db, None


# ------------------------------------------------------------------------------------------------
# norm + FiLM + activation
# ------------------------------------------------------------------------------------------------
@dataclass
class NormCfg:
    mode: int = L.NORM_INSTANCE
    act: int = L.ACT_NONE
    eps: float = 1e-5
    stats: Optional[torch.Tensor] = None          # [B, chunks, C', 2] partial sums from the producing conv
    running_mean: Optional[torch.Tensor] = None
    running_var: Optional[torch.Tensor] = None
    momentum: float = 0.1
    n_updates: int = 1
    update_running: bool = False
    out: Optional[torch.Tensor] = None            # write target (no-grad mode only)
    extra: dict = field(default_factory=dict)


def norm_coefficients(x, g, h, cfg: NormCfg):
    """stats (if not supplied) + finalize -> (A, S, mean, rstd), each [B, C] fp32."""
    B, D, H, W, Cn = x.shape
    V = D * H * W
    dev = x.device
    fa = L.NormFinalizeArgs()
    part = None
    if cfg.mode in (L.NORM_INSTANCE, L.NORM_BATCH):
        part = cfg.stats
        if part is None:
            chunks = L.lib().coma_norm_stats_chunks(V)
            part = torch.empty(B, chunks, Cn, 2, device=dev, dtype=torch.float32)
            L.call("coma_norm_stats", L.ptr(x), B, V, Cn, vol_cs(x), 0, L.dtype_code(x.dtype), L.ptr(part), L.stream())
        elif part.shape[2] != Cn:   # conv computed more (padding) channels than it stored
            part = part[:, :, :Cn, :].contiguous()
        fa.partial, fa.chunks = L.ptr(part), part.shape[1]
    out = torch.empty(4, B, Cn, device=dev, dtype=torch.float32)
    fa.B, fa.C, fa.V, fa.mode, fa.eps = B, Cn, V, cfg.mode, cfg.eps
    if cfg.mode == L.NORM_GIVEN:
        fa.given_mean, fa.given_var = L.ptr(cfg.running_mean), L.ptr(cfg.running_var)
    g = None if g is None else g.detach().float().expand(B, Cn).contiguous()
    h = None if h is None else h.detach().float().expand(B, Cn).contiguous()
    fa.g, fa.h = L.ptr(g), L.ptr(h)
    fa.A, fa.S, fa.mean, fa.rstd = (L.ptr(out[i]) for i in range(4))
    if cfg.mode == L.NORM_BATCH and cfg.update_running and cfg.running_mean is not None:
        fa.running_mean, fa.running_var = L.ptr(cfg.running_mean), L.ptr(cfg.running_var)
        fa.momentum, fa.n_updates = cfg.momentum, cfg.n_updates
    L.call("coma_norm_stats_finalize", C.byref(fa), L.stream())
    return out[0], out[1], out[2], out[3], g


def affine_act(x, A, S, slope, act, out=None, residual=None):
    B, D, H, W, Cn = x.shape
    y = out if out is not None else torch.empty(B, D, H, W, Cn, device=x.device, dtype=x.dtype)
    a = L.AffineActArgs()
    a.x, a.y, a.A, a.S, a.slope = L.ptr(x), L.ptr(y), L.ptr(A), L.ptr(S), L.ptr(slope)
    a.B, a.C, a.V = B, Cn, D * H * W
    a.x_cs, a.x_co, a.y_cs, a.y_co, a.act, a.dtype = vol_cs(x), 0, vol_cs(y), 0, act, L.dtype_code(x.dtype)
    if residual is not None:
        a.r, a.r_cs = L.ptr(residual), vol_cs(residual)
    L.call("coma_norm_film_act_fwd", C.byref(a), L.stream())
    return y


class NormActFn(torch.autograd.Function):
    """y = act(g * (x - mean) * rstd + h [+ residual]) with mean/rstd per cfg.mode; g, h are [B,C] (or None)."""

    @staticmethod
    def forward(ctx, x, g, h, slope, residual, cfg: NormCfg):
        x = as_vol(x)
        residual = None if residual is None else as_vol(residual)
        A, S, mean, rstd, gd = norm_coefficients(x, g, h, cfg)
        sl = None if slope is None else slope.detach().float().reshape(-1)[:1].contiguous()
        y = affine_act(x, A, S, sl, cfg.act, out=cfg.out, residual=residual)
        if cfg.out is not None:
            y = alias(y)               # the caller's slice of a wider (concat) buffer: not a view in autograd's eyes
        ctx.save_for_backward(x, A, S, mean, rstd, gd, sl, residual)
        ctx.cfg = cfg
        ctx.g_shape = None if g is None else g.shape
        ctx.h_shape = None if h is None else h.shape
        ctx.slope_shape = None if slope is None else slope.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, A, S, mean, rstd, gd, sl, residual = ctx.saved_tensors
        cfg: NormCfg = ctx.cfg
        dy = as_vol(dy if dy.dtype == x.dtype else dy.to(x.dtype))
        B, D, H, W, Cn = x.shape
        V = D * H * W
        dev = x.device
        chunks = L.lib().coma_norm_stats_chunks(V)
        partial = torch.empty(B, chunks, Cn, 3, device=dev, dtype=torch.float32)
        dgh = torch.empty(2, B, Cn, device=dev, dtype=torch.float32)
        coef = torch.empty(B, Cn, 3, device=dev, dtype=torch.float32)
        dslope = torch.zeros(1, device=dev, dtype=torch.float32)
        dx = torch.empty(B, D, H, W, Cn, device=dev, dtype=x.dtype)
        a = L.AffineActBwdArgs()
        a.x, a.dy, a.dx = L.ptr(x), L.ptr(dy), L.ptr(dx)
        a.A, a.S, a.mean, a.rstd, a.g, a.slope = L.ptr(A), L.ptr(S), L.ptr(mean), L.ptr(rstd), L.ptr(gd), L.ptr(sl)
        a.B, a.C, a.V = B, Cn, V
        a.x_cs, a.x_co, a.dy_cs, a.dy_co, a.dx_cs, a.dx_co = vol_cs(x), 0, vol_cs(dy), 0, vol_cs(dx), 0
        a.act, a.mode, a.dtype = cfg.act, cfg.mode, L.dtype_code(x.dtype)
        a.partial, a.dg, a.dh, a.dslope, a.coef = L.ptr(partial), L.ptr(dgh[0]), L.ptr(dgh[1]), L.ptr(dslope), L.ptr(coef)
        dr = None
        if residual is not None:
            a.r, a.r_cs = L.ptr(residual), vol_cs(residual)
            if ctx.needs_input_grad[4]:
                dr = torch.empty(B, D, H, W, Cn, device=dev, dtype=x.dtype)
                a.dr, a.dr_cs = L.ptr(dr), Cn
        L.call("coma_norm_film_act_bwd", C.byref(a), L.stream())
        dg = dh = ds = None
        if ctx.g_shape is not None and ctx.needs_input_grad[1]:
            dg = dgh[0].sum_to_size(ctx.g_shape) if tuple(ctx.g_shape) != (B, Cn) else dgh[0]
        if ctx.h_shape is not None and ctx.needs_input_grad[2]:
            dh = dgh[1].sum_to_size(ctx.h_shape) if tuple(ctx.h_shape) != (B, Cn) else dgh[1]
        if ctx.slope_shape is not None and ctx.needs_input_grad[3]:
            ds = dslope.reshape(ctx.slope_shape)
        return (dx if ctx.needs_input_grad[0] else None), dg, dh, ds, dr, None


def norm_act(x, g, h, slope, cfg: NormCfg, residual=None):
    return NormActFn.apply(x, g, h, slope, residual, cfg)


def concat2(a, b):
    """torch.cat((a, b), channel) as two strided identity sweeps of the apply kernel (generic fallback)."""
    return Concat2Fn.apply(a, b)


def alias(t: torch.Tensor) -> torch.Tensor:
    """A new tensor object over the same memory that autograd does not regard as a view of anything: what a custom Function
    returns when its kernel wrote straight into a caller-provided slice of a wider buffer."""
    return torch.empty(0, device=t.device, dtype=t.dtype).set_(t.untyped_storage(), t.storage_offset(), t.size(), t.stride())


class JoinFn(torch.autograd.Function):
    """torch.cat((a, b), channel) WITHOUT a copy (attn_unet_data_parallel.py:229): ``a`` and ``b`` already are the two
    channel halves of ``buf`` (their producers wrote them there); forward hands out ``buf``, backward hands each producer its
    half of the gradient as a channel-sliced view (every backward kernel takes a channel stride)."""

    @staticmethod
    def forward(ctx, a, b, buf):
        Ca, Cb = a.shape[-1], b.shape[-1]
        assert buf.shape[-1] == Ca + Cb and a.data_ptr() == buf.data_ptr() and b.data_ptr() == buf[..., Ca:].data_ptr()
        ctx.split = (Ca, Cb)
        return alias(buf)

    @staticmethod
    def backward(ctx, d):
        Ca, Cb = ctx.split
        d = as_vol(d)
        return (d[..., :Ca] if ctx.needs_input_grad[0] else None), (d[..., Ca:] if ctx.needs_input_grad[1] else None), None


FUSED_FILM = os.environ.get("COMA_DISABLE_FILM_FUSED", "0") != "1"


class FilmAllFn(torch.autograd.Function):
    """Every FiLM MLP of a model (Linear(n, 64) -> ReLU -> Linear(64, 2C) per conditioned layer, specification: DESIGN.md section 4) in ONE
    launch forward and ONE backward (coma_film_mlp_fwd / _bwd) instead of ~12 framework launches per layer and training step.

    ``FilmAllFn.apply(cov, meta, W1_0, b1_0, W2_0, b2_0, W1_1, ...)`` with cov ``[B, n]`` fp32 and meta = ((n_cov_l, C_l), ...)
    returns (dgamma_0, beta_0, dgamma_1, ...), each ``[B, C_l]`` fp32 and contiguous."""

    @staticmethod
    def _args(cov, meta, params):
        a = L.FilmArgs()
        a.n_layers, a.B, a.cov_stride, a.cov = len(meta), cov.shape[0], cov.stride(0), L.ptr(cov)
        for l, (n, Cn) in enumerate(meta):
            a.n_cov[l], a.C[l] = n, Cn
            a.W1[l], a.b1[l], a.W2[l], a.b2[l] = (L.ptr(p) for p in params[4 * l:4 * l + 4])
        return a

    @staticmethod
    def forward(ctx, cov, meta, *params):
        assert cov.dim() == 2 and cov.dtype == torch.float32 and cov.stride(1) == 1 and len(params) == 4 * len(meta)
        params = [p.detach() if p.is_contiguous() else p.detach().contiguous() for p in params]
        B = cov.shape[0]
        flat = torch.empty(sum(2 * B * Cn for _, Cn in meta), device=cov.device, dtype=torch.float32)
        hid = torch.empty(len(meta), B, 64, device=cov.device, dtype=torch.float32)
        a = FilmAllFn._args(cov, meta, params)
        a.hid = L.ptr(hid)
        outs, off = [], 0
        for l, (_, Cn) in enumerate(meta):
            o = flat[off:off + 2 * B * Cn]
            a.out[l] = L.ptr(o)
            outs += [alias(o[:B * Cn].view(B, Cn)), alias(o[B * Cn:].view(B, Cn))]
            off += 2 * B * Cn
        L.call("coma_film_mlp_fwd", C.byref(a), L.stream())
        ctx.save_for_backward(cov, hid, *params)
        ctx.meta = meta
        ctx.set_materialize_grads(False)       # backward must see None for the outputs nobody consumed
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        cov, hid, *params = ctx.saved_tensors
        meta = ctx.meta
        a = FilmAllFn._args(cov, meta, params)
        a.hid = L.ptr(hid)
        gflat = torch.empty(sum(p.numel() for p in params), device=cov.device, dtype=torch.float32)
        pgrads, off, keep = [], 0, []
        for l in range(len(meta)):
            for k, name in enumerate(("dW1", "db1", "dW2", "db2")):
                p = params[4 * l + k]
                g = alias(gflat[off:off + p.numel()].view(p.shape))
                getattr(a, name)[l] = L.ptr(g)
                pgrads.append(g)
                off += p.numel()
            for k, name in enumerate(("d_dgamma", "d_beta")):
                g = grads[2 * l + k]
                if g is not None:
                    g = g.float().contiguous()
                    keep.append(g)
                getattr(a, name)[l] = L.ptr(g)
        L.call("coma_film_mlp_bwd", C.byref(a), L.stream())
        out = []
        for l in range(len(meta)):
            used = grads[2 * l] is not None or grads[2 * l + 1] is not None       # a layer the forward never consumed keeps grad None
            out += [pgrads[4 * l + k] if used and ctx.needs_input_grad[2 + 4 * l + k] else None for k in range(4)]
        return (None, None, *out)


class ForkFn(torch.autograd.Function):
    """t -> (t, t) for a tensor with two consumers one of which hands back a channel-sliced gradient (ops.JoinFn): backward sums the
    two gradients in ONE strided-aware streaming pass (coma_norm_film_act_fwd with identity coefficients and a residual).
    Autograd's own accumulation takes the non-vectorised elementwise kernel for a strided operand -- 3 x the HBM time."""

    @staticmethod
    def forward(ctx, t):
        return alias(t), alias(t)

    @staticmethod
    def backward(ctx, da, db):
        if da is None or db is None:
            return da if db is None else db
        da, db = as_vol(da), as_vol(db)
        if db.dtype != da.dtype:
            db = db.to(da.dtype)
        one, zero = _identity_coefficients(da)
        return affine_act(da, one, zero, None, L.ACT_NONE, residual=db)


class Concat2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = as_vol(a), as_vol(b)
        B, D, H, W, Ca = a.shape
        Cb = b.shape[-1]
        out = torch.empty(B, D, H, W, Ca + Cb, device=a.device, dtype=a.dtype)
        _copy_channels(a, out[..., :Ca])
        _copy_channels(b, out[..., Ca:])
        ctx.split = (Ca, Cb)
        return out

    @staticmethod
    def backward(ctx, d):
        Ca, Cb = ctx.split
        d = as_vol(d)
        da = _copy_channels(d[..., :Ca], None) if ctx.needs_input_grad[0] else None
        db = _copy_channels(d[..., Ca:], None) if ctx.needs_input_grad[1] else None
        return da, db


_ident_cache: dict = {}


def _identity_coefficients(src):
    B, Cn = src.shape[0], src.shape[-1]
    key = (src.device, B, Cn)
    if key not in _ident_cache:
        _ident_cache[key] = (torch.ones(B, Cn, device=src.device), torch.zeros(B, Cn, device=src.device))
    return _ident_cache[key]


def _copy_channels(src, dst):
    one, zero = _identity_coefficients(src)
    return affine_act(src, one, zero, None, L.ACT_NONE, out=dst)


# ------------------------------------------------------------------------------------------------
# attention gate pieces
# ------------------------------------------------------------------------------------------------
def gate_fused(g, x, wg, wx, bsum, wpsi, bpsi, out=None, psi_out=None):
    """Eval-mode gate in one kernel.  wg/wx: [F,C] fp32 (BN folded), bsum/wpsi: [F], bpsi: 1-element tensor."""
    g, x = as_vol(g), as_vol(x)
    B, D, H, W, Cn = x.shape
    o = out if out is not None else torch.empty(B, D, H, W, Cn, device=x.device, dtype=x.dtype)
    a = L.GateArgs()
    a.g, a.x, a.out, a.psi_out = L.ptr(g), L.ptr(x), L.ptr(o), L.ptr(psi_out)
    a.wg, a.wx, a.bsum, a.wpsi = L.ptr(wg), L.ptr(wx), L.ptr(bsum), L.ptr(wpsi)
    a.bpsi, a.bpsi_ptr = 0.0, L.ptr(bpsi)
    a.B, a.C, a.F, a.V = B, Cn, Cn // 2, D * H * W
    a.g_cs, a.g_co, a.x_cs, a.x_co, a.out_cs, a.out_co = vol_cs(g), 0, vol_cs(x), 0, vol_cs(o), 0
    a.dtype = L.dtype_code(x.dtype)
    L.call("coma_gate_fwd", C.byref(a), L.stream())
    return o


class BcastMulFn(torch.autograd.Function):
    """out[b,v,c] = x[b,v,c] * p[b,v,0]"""

    @staticmethod
    def forward(ctx, x, p, out):
        x, p = as_vol(x), p.contiguous()
        B, D, H, W, Cn = x.shape
        o = out if out is not None else torch.empty(B, D, H, W, Cn, device=x.device, dtype=x.dtype)
        a = L.BcastMulArgs()
        a.x, a.p, a.out = L.ptr(x), L.ptr(p), L.ptr(o)
        a.B, a.C, a.V = B, Cn, D * H * W
        a.x_cs, a.x_co, a.out_cs, a.out_co, a.dtype = vol_cs(x), 0, vol_cs(o), 0, L.dtype_code(x.dtype)
        L.call("coma_gate_apply_fwd", C.byref(a), L.stream())
        ctx.save_for_backward(x, p)
        return alias(o) if out is not None else o

    @staticmethod
    def backward(ctx, dout):
        x, p = ctx.saved_tensors
        dout = as_vol(dout if dout.dtype == x.dtype else dout.to(x.dtype))
        B, D, H, W, Cn = x.shape
        dx = torch.empty(B, D, H, W, Cn, device=x.device, dtype=x.dtype)
        dp = torch.empty_like(p)
        a = L.BcastMulArgs()
        a.x, a.p, a.dout, a.dx, a.dp = L.ptr(x), L.ptr(p), L.ptr(dout), L.ptr(dx), L.ptr(dp)
        a.B, a.C, a.V = B, Cn, D * H * W
        a.x_cs, a.x_co, a.out_cs, a.out_co, a.dtype = vol_cs(x), 0, vol_cs(dout), 0, L.dtype_code(x.dtype)
        L.call("coma_gate_bwd", C.byref(a), L.stream())
        return dx, dp, None


def bcast_mul(x, p, out=None):
    return BcastMulFn.apply(x, p, out)


# ------------------------------------------------------------------------------------------------
# ROI painting / packing / loss
# ------------------------------------------------------------------------------------------------
class RoiPaintFn(torch.autograd.Function):
    """[prompt(pos|neg), saliency, suvr, 0...] buffer; gradients flow to the two prompt parameters only."""

    @staticmethod
    def forward(ctx, pos_prompt, neg_prompt, roi, mri, lut, roi_ids, is_pos, cs, dtype, used=(True, True)):
        B = roi.shape[0]
        D, H, W = roi.shape[-3:]
        V = D * H * W
        out = torch.empty(B, D, H, W, cs, device=roi.device, dtype=dtype)
        a = L.RoiPaintArgs()
        roi_f, mri_f = roi.reshape(B, V).float().contiguous(), mri.reshape(B, V).float().contiguous()
        pp, npp = pos_prompt.detach().reshape(-1).float().contiguous(), neg_prompt.detach().reshape(-1).float().contiguous()
        assert pp.numel() == V, f"prompt has {pp.numel()} voxels, input volume has {V} (pass prompt_shape=...)"
        a.roi, a.mri, a.lut, a.roi_ids, a.is_pos = L.ptr(roi_f), L.ptr(mri_f), L.ptr(lut), L.ptr(roi_ids), L.ptr(is_pos)
        a.pos_prompt, a.neg_prompt, a.out = L.ptr(pp), L.ptr(npp), L.ptr(out)
        a.B, a.n_roi, a.out_cs, a.V, a.dtype = B, roi_ids.numel(), cs, V, L.dtype_code(dtype)
        L.call("coma_roi_paint", C.byref(a), L.stream())
        ctx.save_for_backward(is_pos)
        ctx.pshape, ctx.used = pos_prompt.shape, used
        return out

    @staticmethod
    def backward(ctx, dout):
        (is_pos,) = ctx.saved_tensors
        dout = as_vol(dout)
        B, D, H, W, cs = dout.shape
        V = D * H * W
        d = torch.empty(2, V, device=dout.device, dtype=torch.float32)
        L.call("coma_roi_paint_bwd", L.ptr(dout), L.ptr(is_pos), L.ptr(d[0]), L.ptr(d[1]), B, V, vol_cs(dout),
               L.dtype_code(dout.dtype), L.stream())
        # a prompt no sample selected keeps grad None (reference: attn_unet_data_parallel.py:639; SURVEY hard part 5)
        used = ctx.used() if callable(ctx.used) else ctx.used
        dpos = d[0].reshape(ctx.pshape) if used[0] else None
        dneg = d[1].reshape(ctx.pshape) if used[1] else None
        return dpos, dneg, None, None, None, None, None, None, None, None


class Pack2Fn(torch.autograd.Function):
    """dst[...,0] = a + a_add (broadcast over batch), dst[...,1] = b, rest 0."""

    @staticmethod
    def forward(ctx, a_t, a_add, b_t, cs):
        a_t, b_t = a_t.contiguous(), (None if b_t is None else b_t.contiguous())
        B, D, H, W, _ = a_t.shape
        V = D * H * W
        dst = torch.empty(B, D, H, W, cs, device=a_t.device, dtype=a_t.dtype)
        add = None if a_add is None else a_add.detach().reshape(-1).float().contiguous()
        a = L.Pack2Args()
        a.a, a.a_add, a.b, a.dst = L.ptr(a_t), L.ptr(add), L.ptr(b_t), L.ptr(dst)
        a.B, a.dst_cs, a.V, a.dtype = B, cs, V, L.dtype_code(a_t.dtype)
        L.call("coma_pack2_fwd", C.byref(a), L.stream())
        ctx.add_shape = None if a_add is None else a_add.shape
        ctx.shape = a_t.shape
        return dst

    @staticmethod
    def backward(ctx, ddst):
        ddst = ddst.contiguous()
        B, D, H, W, cs = ddst.shape
        V = D * H * W
        da = torch.empty(ctx.shape, device=ddst.device, dtype=ddst.dtype) if ctx.needs_input_grad[0] else None
        db = torch.empty(ctx.shape, device=ddst.device, dtype=ddst.dtype) if ctx.needs_input_grad[2] else None
        if da is None and db is None and ctx.add_shape is None:
            return None, None, None, None
        dadd = torch.empty(V, device=ddst.device, dtype=torch.float32) if ctx.add_shape is not None else None
        a = L.Unpack2Args()
        a.ddst, a.da, a.db, a.d_a_add = L.ptr(ddst), L.ptr(da), L.ptr(db), L.ptr(dadd)
        a.B, a.dst_cs, a.V, a.dtype = B, cs, V, L.dtype_code(ddst.dtype)
        L.call("coma_pack2_bwd", C.byref(a), L.stream())
        return da, (None if dadd is None else dadd.reshape(ctx.add_shape)), db, None


class RoiMseFn(torch.autograd.Function):
    """loss[b] = mean_v(mask_b) * mean_v((pred-gt)^2)   (criterions.py:181-211, voxel_wise=False)."""

    @staticmethod
    def forward(ctx, pred, gt, roi, roi_ids, roi_w):
        B = pred.shape[0]
        V = pred[0].numel()
        p = pred.reshape(B, V).contiguous()
        g, r = gt.reshape(B, V).float().contiguous(), roi.reshape(B, V).float().contiguous()
        chunks = L.lib().coma_roi_mse_chunks(V)
        partial = torch.empty(B, chunks, 2, device=p.device, dtype=torch.float32)
        loss = torch.empty(B, device=p.device, dtype=torch.float32)
        sums = torch.empty(B, 2, device=p.device, dtype=torch.float32)
        a = L.RoiMseArgs()
        a.pred, a.gt, a.roi, a.roi_ids, a.roi_w, a.n_roi = L.ptr(p), L.ptr(g), L.ptr(r), L.ptr(roi_ids), L.ptr(roi_w), roi_ids.numel()
        a.B, a.V, a.dtype = B, V, L.dtype_code(p.dtype)
        a.partial, a.loss, a.sums = L.ptr(partial), L.ptr(loss), L.ptr(sums)
        L.call("coma_roi_mse_fwd", C.byref(a), L.stream())
        ctx.save_for_backward(p, g, sums)
        ctx.pshape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        p, g, sums = ctx.saved_tensors
        B, V = p.shape
        dpred = torch.empty_like(p)
        dl = dloss.float().contiguous()
        a = L.RoiMseArgs()
        a.pred, a.gt, a.sums, a.dloss, a.dpred = L.ptr(p), L.ptr(g), L.ptr(sums), L.ptr(dl), L.ptr(dpred)
        a.B, a.V, a.dtype = B, V, L.dtype_code(p.dtype)
        L.call("coma_roi_mse_bwd", C.byref(a), L.stream())
        return dpred.reshape(ctx.pshape), None, None, None, None
