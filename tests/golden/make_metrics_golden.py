"""Golden vectors for the evaluation metrics, produced by the REFERENCE's own code on CPU:
``calc_roi_metrics`` (attn_unet_data_parallel.py:1361-1397) is called as is (its ``torch.zeros(device=roi.get_device())``
idiom needs the device -1 -> "cpu" proxy of make_golden), and the batch metric lines :1214-1231 are executed verbatim below.

    python -m tests.golden.make_metrics_golden         # build container only; writes tests/golden/metrics_golden.npz
"""
from __future__ import annotations

import os

import numpy as np
import torch

from tests.golden import common, make_golden

HERE = os.path.dirname(os.path.abspath(__file__))


def case(batch, shape, seed):
    mri, tau, roi, covars, dicts = common.synthetic_batch(batch, shape, seed)
    g = torch.Generator().manual_seed(seed + 77)
    pred = (tau + 0.2 * torch.randn(tau.shape, generator=g)).clamp_min(0) * (mri > 0)     # a plausible prediction, zero outside the brain
    return pred, tau, roi


def main():
    ref_model, ref_crit = make_golden.import_reference()
    ref_model.torch = ref_crit.torch                # the same device -1 -> cpu proxy for calc_roi_metrics
    out = {}
    for name, (batch, shape, seed) in {"m32": (3, (32, 32, 32), 41), "m48": (2, (48, 40, 56), 42)}.items():
        pred, tau_volume, roi = case(batch, shape, seed)
        # ---- attn_unet_data_parallel.py:1214-1231, verbatim ----
        diff = pred - tau_volume
        mae = torch.mean(torch.abs(diff))
        raw_mape = torch.abs(diff / tau_volume)
        nr_mape = torch.where(torch.abs(tau_volume) > 1e-08, torch.abs((tau_volume - pred) / tau_volume), torch.nan)
        mape = torch.nansum(nr_mape * 100, dim=(-3, -2, -1)).sum()
        gt_mean = torch.mean(tau_volume, dim=(-3, -2, -1))
        squared_error_num = torch.sum(torch.square(tau_volume - pred), dim=(-3, -2, -1))
        squared_error_den = torch.sum(torch.square(tau_volume - gt_mean.view(-1, 1, 1, 1, 1)), dim=(-3, -2, -1))
        rse = torch.mean(squared_error_num / squared_error_den)
        num = torch.sum(torch.square(tau_volume - pred), dim=(-3, -2, -1))
        den = torch.sum(torch.square(tau_volume), dim=(-3, -2, -1))
        rrmse = torch.nanmean(torch.sqrt(num / den))
        n = len(common.ROI_INDICES)
        z = [torch.zeros(n) for _ in range(5)]
        r = ref_model.calc_roi_metrics(common.ROI_INDICES, None, *z, tau_volume, roi, pred, diff, raw_mape)
        out[f"{name}/volume"] = np.array([float(mae), float(mape), float(rse), float(rrmse)], dtype=np.float64)
        for key, t in zip(("roi_maes", "roi_mapes", "roi_rses", "roi_wrrmses", "roi_nonnan"), r):
            out[f"{name}/{key}"] = t.double().numpy()
        out[f"{name}/cfg"] = np.array([batch, *shape, seed])
    np.savez_compressed(os.path.join(HERE, "metrics_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
