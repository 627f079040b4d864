"""Evaluation metrics of the step right after inference, on one fused CUDA pass.

The reference computes MAE / MAPE / RSE / RRMSE per batch (attn_unet_data_parallel.py:1214-1231) and, in
``calc_roi_metrics`` (:1361-1397), 36 ROIs x ~10 masked full-volume kernels.  ``coma_eval_metrics`` reads pred / tau / roi
once and accumulates eight fp64 sums per (sample, ROI slot) and per (sample, all voxels); everything the reference returns
is a closed form of those sums, evaluated here on a [B, 37, 8] tensor.  Same function names, argument order and return
values as the reference.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_ids_cache: dict = {}


def fused_sums(pred, tau_volume, roi, roi_indices):
    """-> float64 ``[B, len(roi_indices) + 1, 8]``: voxels, sum|d|, sum d^2, sum tau, sum tau^2, nansum|d/tau|, NaN count,
    sum 100|d/tau| over |tau| > 1e-8 (last slot = all voxels)."""
    B = pred.shape[0]
    p, t, r = (x.reshape(B, -1).float().contiguous() for x in (pred, tau_volume, roi))
    key = (tuple(int(i) for i in roi_indices), str(p.device))
    ids = _ids_cache.get(key)
    if ids is None:
        ids = _ids_cache[key] = torch.tensor(key[0], dtype=torch.int32, device=p.device)
    out = torch.zeros(B, len(key[0]) + 1, 8, dtype=torch.float64, device=p.device)
    a = L.EvalMetricsArgs()
    a.pred, a.tau, a.roi, a.roi_ids, a.out = L.ptr(p), L.ptr(t), L.ptr(r), L.ptr(ids), L.ptr(out)
    a.n_roi, a.B, a.V = len(key[0]), B, p.shape[1]
    L.call("coma_eval_metrics", C.byref(a), L.stream())
    return out


def volume_metrics(pred, tau_volume, sums=None, roi=None, roi_indices=()):
    """(mae, mape, rse, rrmse) increments of one batch, attn_unet_data_parallel.py:1214-1231."""
    if sums is None:
        sums = fused_sums(pred, tau_volume, roi if roi is not None else torch.zeros_like(tau_volume), roi_indices)
    g = sums[:, -1, :]
    n = g[:, 0]
    mae = g[:, 1].sum() / n.sum()
    mape = g[:, 7].sum()
    rse = (g[:, 2] / (g[:, 4] - g[:, 3] ** 2 / n)).mean()
    rrmse = torch.nanmean(torch.sqrt(g[:, 2] / g[:, 4]))
    return mae.float(), mape.float(), rse.float(), rrmse.float()


def calc_roi_metrics(roi_indices, roi_weights, roi_maes, roi_mapes, roi_rses, roi_wrrmses, roi_nonnan_voxels, tau_volume, roi,
                     pred, diff=None, raw_mape=None, sums=None):
    """Drop-in for attn_unet_data_parallel.py:1361-1397 (same signature; the accumulator arguments, ``diff`` and ``raw_mape``
    are accepted and ignored like the reference ignores its accumulators).  -> (roi_maes, roi_mapes, roi_rses, roi_wrrmses,
    roi_nonnan_voxels), each ``[len(roi_indices)]`` float32 on the device."""
    if sums is None:
        sums = fused_sums(pred, tau_volume, roi, roi_indices)
    s = sums[:, :-1, :]                                  # [B, R, 8]
    n = s[..., 0]
    maes = (s[..., 1] / n).sum(dim=0)
    mapes = s[..., 5].sum(dim=0)
    nonnan = (n - s[..., 6]).sum(dim=0)
    wrrmses = torch.sqrt(s[..., 2] / s[..., 4]).sum(dim=0)
    rses = (s[..., 2] / (s[..., 4] - s[..., 3] ** 2 / n)).sum(dim=0)
    return maes.float(), mapes.float(), rses.float(), wrrmses.float(), nonnan.float()
