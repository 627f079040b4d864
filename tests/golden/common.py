"""Shared, machine-independent generators for the golden fixtures (weights and inputs).

Everything is derived from CPU ``torch.Generator`` streams seeded by (seed, crc32(name)), so the
GPU box regenerates bit-identical weights/inputs from the few integers stored in the fixtures
instead of shipping 100+ MB state dicts.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch
import torch.nn as nn

ROI_INDICES = [
    1001, 1006, 1007, 1009, 1015, 1016, 1030, 1034, 1033, 1008, 1025, 1029, 1031, 1022, 17, 18,
    2001, 2006, 2007, 2009, 2015, 2016, 2030, 2034, 2033, 2008, 2025, 2029, 2031, 2022, 49, 50, 51, 52, 53, 54,
]


def _gen(seed: int, name: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))


def _randn(shape, seed, name):
    return torch.randn(tuple(shape), generator=_gen(seed, name), dtype=torch.float32)


def _rand(shape, seed, name):
    return torch.rand(tuple(shape), generator=_gen(seed, name), dtype=torch.float32)


@torch.no_grad()
def fill_deterministic(model: nn.Module, seed: int = 0) -> nn.Module:
    """Overwrite every parameter/buffer with a value that depends only on (seed, key, shape)."""
    kinds = {}
    for mod_name, mod in model.named_modules():
        prefix = mod_name + "." if mod_name else ""
        if isinstance(mod, nn.modules.batchnorm._BatchNorm):
            kinds[prefix + "weight"], kinds[prefix + "bias"] = "bn_w", "small"
            kinds[prefix + "running_mean"], kinds[prefix + "running_var"] = "small", "var"
            kinds[prefix + "num_batches_tracked"] = "zero"
        elif isinstance(mod, nn.PReLU):
            kinds[prefix + "weight"] = "prelu"
        elif isinstance(mod, nn.ConvTranspose3d):   # weight is [Cin, Cout, k, k, k]; stride 2 -> 27/8 taps per output
            kinds[prefix + "weight"], kinds[prefix + "bias"] = "fan_t", "small"
        elif isinstance(mod, (nn.Conv3d, nn.Linear)):
            kinds[prefix + "weight"], kinds[prefix + "bias"] = "fan", "small"
    for key, t in model.state_dict().items():
        kind = kinds.get(key)
        if kind is None:
            if key.endswith("prompt"):
                kind = "randn"
            elif "reweigh" in key:
                kind = "one"
            elif key.endswith(".bias"):
                kind = "small"
            elif t.dim() >= 2:
                kind = "fan"  # expert-mixed kernels [E, Cout, Cin, k, k, k]
            else:
                kind = "small"
        if kind == "zero":
            t.zero_()
        elif kind == "one":
            t.fill_(1.0)
        elif kind == "randn":
            t.copy_(_randn(t.shape, seed, key))
        elif kind == "small":
            t.copy_(0.1 * _randn(t.shape, seed, key))
        elif kind == "var":
            t.copy_(1.0 + 0.2 * _rand(t.shape, seed, key))
        elif kind == "bn_w":
            t.copy_(1.0 + 0.1 * _randn(t.shape, seed, key))
        elif kind == "prelu":
            t.copy_(0.25 + 0.05 * _randn(t.shape, seed, key))
        elif kind in ("fan", "fan_t"):
            if kind == "fan_t":
                fan_in = t.shape[0] * t[0, 0].numel() / 8.0
            elif t.dim() == 6:        # experts
                fan_in = t[0, 0].numel()
            else:
                fan_in = t[0].numel()
            scale = 1.0 / np.sqrt(max(fan_in, 1))
            if ".film.2." in key:
                scale = 0.3 / 8.0     # non-zero so the FiLM modulation is exercised
            t.copy_(scale * _randn(t.shape, seed, key))
        else:
            raise AssertionError(kind)
    return model


def synthetic_batch(batch: int, shape, seed: int = 1234, covar_dtype=torch.float64):
    """Synthetic (mri, tau, roi, covars, roi_pred_dicts) with the VolumeDataset tuple layout."""
    from_names = roi_names()
    D, H, W = shape
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, D), torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    brain = ((zz / 0.9) ** 2 + (yy / 0.85) ** 2 + (xx / 0.8) ** 2) < 1.0
    mri = _rand((batch, 1, D, H, W), seed, "mri") * brain
    tau = (1.0 + 0.3 * _randn((batch, 1, D, H, W), seed, "tau")).clamp(0, 4) * brain
    bs = max(1, min(D, H, W) // 16)   # label blocks
    gd, gh, gw = -(-D // bs), -(-H // bs), -(-W // bs)
    pick = torch.randint(0, 2 * len(ROI_INDICES), (batch, 1, gd, gh, gw), generator=_gen(seed, "roi"))
    table = torch.tensor(ROI_INDICES + [2, 41] * (len(ROI_INDICES) // 2), dtype=torch.float32)
    roi = table[pick]
    roi = roi.repeat_interleave(bs, 2).repeat_interleave(bs, 3).repeat_interleave(bs, 4)[:, :, :D, :H, :W] * brain
    u = _rand((batch, 8), seed, "covars")
    covars = torch.stack([(u[:, 0] < 0.4).float(), u[:, 1], (u[:, 2] < 0.5).float(), u[:, 3], u[:, 4],
                          1.0 + 0.3 * u[:, 5]], dim=1).reshape(batch, 1, 6).to(covar_dtype)
    locs = 1.0 + 0.3 * _rand((batch, len(ROI_INDICES)), seed, "loc")
    stds = 0.1 * _rand((batch, len(ROI_INDICES)), seed, "std")
    dicts = [{name: {"loc": float(locs[b, i]), "std": float(stds[b, i])} for i, name in enumerate(from_names)}
             for b in range(batch)]
    return mri.contiguous(), tau.contiguous(), roi.contiguous(), covars, dicts


def roi_names():
    ctx = ["bankssts", "entorhinal", "fusiform", "inferiortemporal", "middletemporal", "parahippocampal",
           "superiortemporal", "transversetemporal", "temporalpole", "inferiorparietal", "precuneus",
           "superiorparietal", "supramarginal", "postcentral"]
    return ([f"ctx-lh-{n}" for n in ctx] + ["Left-Hippocampus", "Left-Amygdala"] + [f"ctx-rh-{n}" for n in ctx]
            + ["Right-Thalamus-Proper", "Right-Caudate", "Right-Putamen", "Right-Pallidum", "Right-Hippocampus",
               "Right-Amygdala"])
