"""ctypes binding of libcoma_b200.so (include/coma_b200.h).

The product path has no CPU fallback: if the shared library is missing, or a tensor is not on a
CUDA device, calls raise.  Build with ``python -c "import __graft_entry__ as g; g.build()"``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcoma_b200.so")

F32, BF16, BF16_F32OUT = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_LEAKY_RELU = 0, 1, 2, 3, 4
NORM_NONE, NORM_INSTANCE, NORM_BATCH, NORM_GIVEN = 0, 1, 2, 3
IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = 0, 1, 2

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class ConvArgs(C.Structure):
    _fields_ = [("x", _vp), ("w", _vp), ("bias", _vp), ("y", _vp), ("scale", _vp), ("shift", _vp), ("slope", _vp),
                ("stats", _vp),
                ("B", _i32), ("Di", _i32), ("Hi", _i32), ("Wi", _i32), ("Do", _i32), ("Ho", _i32), ("Wo", _i32),
                ("Cin", _i32), ("Cout", _i32), ("x_cs", _i32), ("x_co", _i32), ("y_cs", _i32), ("y_co", _i32),
                ("y_cn", _i32), ("ksize", _i32), ("stride", _i32), ("pad", _i32), ("transposed", _i32),
                ("w_bstride", _i64), ("bias_bstride", _i32), ("act", _i32), ("dtype", _i32), ("impl", _i32),
                ("in_scale", _vp), ("in_shift", _vp), ("in_slope", _vp), ("in_act", _i32), ("reserved0", _i32)]


class EvalMetricsArgs(C.Structure):
    _fields_ = [("pred", _vp), ("tau", _vp), ("roi", _vp), ("roi_ids", _vp), ("n_roi", _i32), ("B", _i32), ("V", _i64),
                ("out", _vp)]


class PrepareArgs(C.Structure):
    _fields_ = [("mri", _vp), ("tau", _vp), ("roi", _vp), ("mri_out", _vp), ("tau_out", _vp), ("roi_out", _vp),
                ("in_size", _i32 * 3), ("res_size", _i32 * 3), ("out_size", _i32 * 3), ("pad_before", _i32 * 3),
                ("ratio", C.c_double * 3), ("default_value", C.c_float)]


class WgradArgs(C.Structure):
    _fields_ = [("g", _vp), ("x", _vp), ("dw", _vp),
                ("B", _i32), ("Dg", _i32), ("Hg", _i32), ("Wg", _i32), ("Dx", _i32), ("Hx", _i32), ("Wx", _i32),
                ("Cg", _i32), ("Cx", _i32), ("g_cs", _i32), ("g_co", _i32), ("x_cs", _i32), ("x_co", _i32),
                ("ksize", _i32), ("stride", _i32), ("pad", _i32), ("dtype", _i32), ("impl", _i32),
                ("workspace", _vp), ("workspace_bytes", _i64)]


class WeightLayoutArgs(C.Structure):
    _fields_ = [("param", _vp), ("packed", _vp), ("A", _i32), ("B", _i32), ("T", _i32), ("R_pad", _i32), ("C_pad", _i32),
                ("swap", _i32), ("flip", _i32), ("dtype", _i32), ("unpack", _i32)]


FILM_MAX_LAYERS = 32


class FilmArgs(C.Structure):
    _fields_ = [("n_layers", _i32), ("B", _i32), ("cov_stride", _i32), ("cov", _vp), ("hid", _vp),
                ("n_cov", _i32 * FILM_MAX_LAYERS), ("C", _i32 * FILM_MAX_LAYERS)] + [
        (name, _vp * FILM_MAX_LAYERS) for name in ("W1", "b1", "W2", "b2", "out", "d_dgamma", "d_beta", "dW1", "db1", "dW2", "db2")]


class NormFinalizeArgs(C.Structure):
    _fields_ = [("partial", _vp), ("chunks", _i32), ("B", _i32), ("C", _i32), ("V", _i64), ("mode", _i32),
                ("given_mean", _vp), ("given_var", _vp), ("eps", _f32), ("g", _vp), ("h", _vp), ("A", _vp), ("S", _vp),
                ("mean", _vp), ("rstd", _vp), ("running_mean", _vp), ("running_var", _vp), ("momentum", _f32),
                ("n_updates", _i32)]


class AffineActArgs(C.Structure):
    _fields_ = [("x", _vp), ("y", _vp), ("A", _vp), ("S", _vp), ("slope", _vp), ("B", _i32), ("C", _i32), ("V", _i64),
                ("x_cs", _i32), ("x_co", _i32), ("y_cs", _i32), ("y_co", _i32), ("act", _i32), ("dtype", _i32),
                ("r", _vp), ("r_cs", _i32)]


class AffineActBwdArgs(C.Structure):
    _fields_ = [("x", _vp), ("dy", _vp), ("dx", _vp), ("A", _vp), ("S", _vp), ("mean", _vp), ("rstd", _vp), ("g", _vp),
                ("slope", _vp), ("B", _i32), ("C", _i32), ("V", _i64),
                ("x_cs", _i32), ("x_co", _i32), ("dy_cs", _i32), ("dy_co", _i32), ("dx_cs", _i32), ("dx_co", _i32),
                ("act", _i32), ("mode", _i32), ("dtype", _i32),
                ("partial", _vp), ("dg", _vp), ("dh", _vp), ("dslope", _vp), ("coef", _vp),
                ("r", _vp), ("r_cs", _i32), ("dr", _vp), ("dr_cs", _i32)]


class GateArgs(C.Structure):
    _fields_ = [("g", _vp), ("x", _vp), ("out", _vp), ("psi_out", _vp), ("wg", _vp), ("wx", _vp), ("bsum", _vp),
                ("wpsi", _vp), ("bpsi", _f32), ("bpsi_ptr", _vp), ("B", _i32), ("C", _i32), ("F", _i32), ("V", _i64),
                ("g_cs", _i32), ("g_co", _i32), ("x_cs", _i32), ("x_co", _i32), ("out_cs", _i32), ("out_co", _i32),
                ("dtype", _i32)]


class BcastMulArgs(C.Structure):
    _fields_ = [("x", _vp), ("p", _vp), ("out", _vp), ("dout", _vp), ("dx", _vp), ("dp", _vp),
                ("B", _i32), ("C", _i32), ("V", _i64), ("x_cs", _i32), ("x_co", _i32), ("out_cs", _i32), ("out_co", _i32),
                ("dtype", _i32)]


class RoiPaintArgs(C.Structure):
    _fields_ = [("roi", _vp), ("mri", _vp), ("lut", _vp), ("roi_ids", _vp), ("is_pos", _vp), ("pos_prompt", _vp),
                ("neg_prompt", _vp), ("out", _vp), ("B", _i32), ("n_roi", _i32), ("out_cs", _i32), ("V", _i64),
                ("dtype", _i32)]


class Pack2Args(C.Structure):
    _fields_ = [("a", _vp), ("a_add", _vp), ("b", _vp), ("dst", _vp), ("B", _i32), ("dst_cs", _i32), ("V", _i64),
                ("dtype", _i32)]


class Unpack2Args(C.Structure):
    _fields_ = [("ddst", _vp), ("da", _vp), ("db", _vp), ("d_a_add", _vp), ("B", _i32), ("dst_cs", _i32), ("V", _i64),
                ("dtype", _i32)]


class RoiMseArgs(C.Structure):
    _fields_ = [("pred", _vp), ("gt", _vp), ("roi", _vp), ("roi_ids", _vp), ("roi_w", _vp), ("n_roi", _i32),
                ("B", _i32), ("V", _i64), ("dtype", _i32), ("partial", _vp), ("loss", _vp), ("sums", _vp),
                ("dloss", _vp), ("dpred", _vp)]


EXPORTS = {
    # name: (restype, argtypes)
    "coma_version": (C.c_int, []),
    "coma_last_error": (C.c_char_p, []),
    "coma_conv3d_stat_chunks": (C.c_int, [C.POINTER(ConvArgs)]),
    "coma_conv3d_tcgen05_supported": (C.c_int, [C.POINTER(ConvArgs)]),
    "coma_conv3d_impl": (C.c_int, [C.POINTER(ConvArgs)]),
    "coma_conv3d_wgrad_tcgen05_supported": (C.c_int, [C.POINTER(WgradArgs)]),
    "coma_conv3d_wgrad_workspace_size": (_i64, [C.POINTER(WgradArgs)]),
    "coma_conv3d_prologue_supported": (C.c_int, [C.POINTER(ConvArgs)]),
    "coma_conv3d_fprop": (C.c_int, [C.POINTER(ConvArgs), _vp]),
    "coma_convT3d_fprop": (C.c_int, [C.POINTER(ConvArgs), _vp]),
    "coma_conv3d_dgrad": (C.c_int, [C.POINTER(ConvArgs), _vp]),
    "coma_convT3d_dgrad": (C.c_int, [C.POINTER(ConvArgs), _vp]),
    "coma_conv3d_wgrad": (C.c_int, [C.POINTER(WgradArgs), _vp]),
    "coma_convT3d_wgrad": (C.c_int, [C.POINTER(WgradArgs), _vp]),
    "coma_weight_layout": (C.c_int, [C.POINTER(WeightLayoutArgs), _vp]),
    "coma_film_mlp_fwd": (C.c_int, [C.POINTER(FilmArgs), _vp]),
    "coma_film_mlp_bwd": (C.c_int, [C.POINTER(FilmArgs), _vp]),
    "coma_norm_stats_chunks": (C.c_int, [_i64]),
    "coma_norm_stats": (C.c_int, [_vp, _i32, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "coma_gate_stats": (C.c_int, [_vp, _i32, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "coma_norm_stats_finalize": (C.c_int, [C.POINTER(NormFinalizeArgs), _vp]),
    "coma_norm_film_act_fwd": (C.c_int, [C.POINTER(AffineActArgs), _vp]),
    "coma_norm_film_act_bwd": (C.c_int, [C.POINTER(AffineActBwdArgs), _vp]),
    "coma_gate_fwd": (C.c_int, [C.POINTER(GateArgs), _vp]),
    "coma_gate_apply_fwd": (C.c_int, [C.POINTER(BcastMulArgs), _vp]),
    "coma_gate_bwd": (C.c_int, [C.POINTER(BcastMulArgs), _vp]),
    "coma_roi_paint": (C.c_int, [C.POINTER(RoiPaintArgs), _vp]),
    "coma_roi_paint_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp]),
    "coma_pack2_fwd": (C.c_int, [C.POINTER(Pack2Args), _vp]),
    "coma_pack2_bwd": (C.c_int, [C.POINTER(Unpack2Args), _vp]),
    "coma_roi_mse_chunks": (C.c_int, [_i64]),
    "coma_roi_mse_fwd": (C.c_int, [C.POINTER(RoiMseArgs), _vp]),
    "coma_roi_mse_bwd": (C.c_int, [C.POINTER(RoiMseArgs), _vp]),
    "coma_eval_metrics": (C.c_int, [C.POINTER(EvalMetricsArgs), _vp]),
    "coma_prepare_volumes": (C.c_int, [C.POINTER(PrepareArgs), _vp]),
    "coma_upload_small": (C.c_int, [_vp, _vp, _i64, _vp]),
}

_lib = None
launches = 0   # kernels-family calls issued through the C ABI (bench.py reports it as gpu_launches)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension is required (no CPU fallback). "
                               "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream() -> int:
    """Raw cudaStream_t of torch's current stream (the private fast accessor when available: this is on the per-launch path)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Invoke an ABI function that returns a status code; raise with the library's message on failure."""
    global launches
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (status {rc}): {lib().coma_last_error().decode()}")
    launches += 1


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  CPU tensors are rejected: the product path is CUDA only."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("coma_unet_b200 runs on CUDA tensors only (no CPU fallback)")
    return t.data_ptr()


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported activation dtype {dt}")
