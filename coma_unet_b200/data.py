"""Synthetic stand-in for the reference datasets with the same ``__getitem__`` contract:

    (mri[1,D,H,W] f32, tau[1,D,H,W] f32, roi[1,D,H,W] f32 FreeSurfer labels, (abeta, covars[1,6]), path)

(VolumeDataset_ADNI_A4_combined.py:86-91, VolumeDataset_Inference.py:145, VolumeDataset.py:382,646-647).
Covariate order ``[Abeta, Age, Sex, Education, Cognition, Tau_Meta]``; ``covars`` is float64 for the
"combined"/"inference" flavours and float32 for the "adni" flavour (VolumeDataset.py:427).  Value
distributions follow SURVEY.md section 8(d).  There is no network and no lab data here, so this is
what tests and bench.py feed the model.
"""
from __future__ import annotations

import torch
from torch.utils.data import Dataset

from .model import ROI_INDICES, ROI_NAMES


class SyntheticVolumeDataset(Dataset):
    def __init__(self, length=8, shape=(128, 128, 128), seed=1234, flavour="combined", device="cpu"):
        self.length, self.shape, self.seed, self.flavour, self.device = length, tuple(shape), seed, flavour, device
        D, H, W = self.shape
        zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, D), torch.linspace(-1, 1, H), torch.linspace(-1, 1, W),
                                    indexing="ij")
        self.brain = ((zz / 0.9) ** 2 + (yy / 0.85) ** 2 + (xx / 0.8) ** 2) < 1.0
        self.table = torch.tensor(ROI_INDICES + [2, 41] * (len(ROI_INDICES) // 2), dtype=torch.float32)

    def __len__(self):
        return self.length

    def roi_predictions(self, index):
        """Per-sample ROI prediction dict (the JSON lookups of attn_unet_data_parallel.py:708-710,809-810)."""
        g = torch.Generator().manual_seed(self.seed * 7919 + index * 2 + 1)
        loc = 1.0 + 0.3 * torch.rand(len(ROI_NAMES), generator=g)
        std = 0.1 * torch.rand(len(ROI_NAMES), generator=g)
        return {n: {"loc": float(loc[i]), "std": float(std[i])} for i, n in enumerate(ROI_NAMES)}

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 7919 + index * 2)
        D, H, W = self.shape
        mri = torch.rand(1, D, H, W, generator=g) * self.brain
        tau = (1.0 + 0.3 * torch.randn(1, D, H, W, generator=g)).clamp(0, 4) * self.brain
        bs = max(1, min(D, H, W) // 16)
        gd, gh, gw = -(-D // bs), -(-H // bs), -(-W // bs)
        pick = torch.randint(0, len(self.table), (1, gd, gh, gw), generator=g)
        roi = self.table[pick].repeat_interleave(bs, 1).repeat_interleave(bs, 2).repeat_interleave(bs, 3)
        roi = roi[:, :D, :H, :W] * self.brain
        u = torch.rand(6, generator=g)
        covars = torch.tensor([[float(u[0] < 0.4), float(u[1]), float(u[2] < 0.5), float(u[3]), float(u[4]),
                                1.0 + 0.3 * float(u[5])]],
                              dtype=torch.float32 if self.flavour == "adni" else torch.float64)
        abeta = covars[0, 0].to(torch.float32)
        dev = self.device
        return (mri.to(dev), tau.to(dev), roi.to(dev), (abeta, covars), f"/synthetic/adni/{index:03d}-S-0000/PET/analysis/suvr.nii")
