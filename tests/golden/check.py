"""Helpers to compare tensors against the sampled fixtures in golden.npz."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    data = np.load(os.path.join(HERE, "golden.npz"))
    with open(os.path.join(HERE, "golden_meta.json")) as f:
        meta = json.load(f)
    return data, meta


def load2():
    """Round-2 fixtures (tests/golden/make_golden2.py): train64, eval128_full."""
    data = np.load(os.path.join(HERE, "golden2.npz"))
    with open(os.path.join(HERE, "golden2_meta.json")) as f:
        meta = json.load(f)
    return data, meta


def rel_err(a: np.ndarray, b: np.ndarray, floor_frac: float = 1e-3) -> float:
    """Per-voxel relative error max |a-b| / max(|b|, floor_frac*max|b|)  (SURVEY.md section 7 hard part 7)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    floor = floor_frac * max(np.abs(b).max(), 1e-30)
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


def scaled_err(a, b) -> float:
    """max |a-b| / max|b| -- for tensors where only the overall scale is meaningful (gradients)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def sampled(data, prefix: str, t: torch.Tensor):
    """Return (values of t at the fixture's sample positions, fixture values)."""
    idx = data[f"{prefix}/idx"]
    assert list(t.shape) == list(data[f"{prefix}/shape"]), (prefix, t.shape, data[f"{prefix}/shape"])
    got = t.detach().float().cpu().reshape(-1).numpy()[idx]
    return got, data[f"{prefix}/val"]
