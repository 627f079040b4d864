"""CPU: the C-ABI library loads and exports every symbol include/coma_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "coma_b200.h")
LIB = os.path.join(ROOT, "coma_unet_b200", "csrc", "libcoma_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(coma_[a-zA-Z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_survey_entry_points():
    names = declared_symbols()
    for required in ["coma_conv3d_fprop", "coma_conv3d_dgrad", "coma_conv3d_wgrad", "coma_convT3d_fprop",
                     "coma_convT3d_dgrad", "coma_convT3d_wgrad", "coma_norm_stats_finalize", "coma_norm_film_act_fwd",
                     "coma_norm_film_act_bwd", "coma_gate_stats", "coma_gate_fwd", "coma_gate_bwd", "coma_roi_paint",
                     "coma_roi_mse_fwd", "coma_roi_mse_bwd", "coma_version", "coma_last_error"]:
        assert required in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header(lib):
    from coma_unet_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    assert _lib.lib().coma_version() >= 100


def test_bad_arguments_return_status_not_crash(lib):
    from coma_unet_b200 import _lib
    L = _lib.lib()
    assert L.coma_conv3d_fprop(None, None) != 0
    assert b"coma_conv3d_fprop" in L.coma_last_error()
    a = _lib.ConvArgs()
    assert L.coma_conv3d_fprop(ctypes.byref(a), None) != 0
    # round-2 entry points: argument checks come before any launch, so they run without a GPU
    w = _lib.WeightLayoutArgs()
    assert L.coma_weight_layout(ctypes.byref(w), None) != 0 and b"coma_weight_layout" in L.coma_last_error()
    buf = (ctypes.c_float * 64)()
    w.param = w.packed = ctypes.cast(buf, ctypes.c_void_p)
    w.A, w.B, w.T, w.R_pad, w.C_pad = 1, 1, 64, 16, 16            # more than 27 taps
    assert L.coma_weight_layout(ctypes.byref(w), None) != 0 and b"taps" in L.coma_last_error()
    f = _lib.FilmArgs()
    assert L.coma_film_mlp_fwd(ctypes.byref(f), None) != 0 and b"coma_film_mlp" in L.coma_last_error()
    f.cov = f.hid = ctypes.cast(buf, ctypes.c_void_p)
    f.n_layers, f.B, f.cov_stride = _lib.FILM_MAX_LAYERS + 1, 4, 6
    assert L.coma_film_mlp_bwd(ctypes.byref(f), None) != 0
    f.n_layers = 1                                                 # a layer without parameters
    assert L.coma_film_mlp_fwd(ctypes.byref(f), None) != 0 and b"layer 0" in L.coma_last_error()
    g = _lib.WgradArgs()
    assert L.coma_conv3d_wgrad_workspace_size(ctypes.byref(g)) == 0
