"""Oracle restatement of the three loss classes on the training step.  Test infrastructure.

Follows criterions.py:124-211 (``RoiMSE``), :485-575 (``GenerativeContrastiveLoss``) and
:579-644 (``LabelDifference`` / ``FeatureSimilarity`` / ``RnCLoss``).  Pinned: the reference file
itself runs on CPU in this container once ``data_util`` / ``VolumeDataset`` are stubbed and the
``device=roi.get_device()`` idiom (:182, raises for CPU tensors) is routed to ``roi.device``;
tests/golden/make_golden.py records its outputs and gradients, tests/test_oracle_golden.py
compares.  Device-agnostic, no logging (the reference's per-batch ``logging.info`` of tensor
norms, :203-204,571-573, are host syncs, not arithmetic).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class RoiMSE(nn.Module):
    def __init__(self, roi_weights, roi_indices, reduction="mean", scale_factor=360, voxel_wise=True):
        super().__init__()
        self.roi_weights, self.roi_indices = roi_weights, roi_indices
        self.batch_reduction, self.scale_factor, self.voxel_wise = reduction, scale_factor, voxel_wise
        if voxel_wise:  # :135-144 needs the lab-private template volume (data_util.load_template)
            raise NotImplementedError("voxel_wise=True is not reachable from the shipped config (validation.py:146)")
        self.voxel_weights = None

    def forward(self, pred, gt, roi):
        mask = torch.zeros(roi.size(), device=roi.device)
        for w, idx in zip(self.roi_weights, self.roi_indices):
            mask[roi == idx] = w
        per_sample = torch.mean(torch.square(pred - gt), dim=(-3, -2, -1))                     # [B,1]
        loss = torch.stack([torch.mean(mask[b] * per_sample[b]) for b in range(pred.size(0))]).reshape(per_sample.shape)
        return torch.mean(loss) if self.batch_reduction == "mean" else loss


class LabelDifference(nn.Module):
    def __init__(self, distance_type="l1"):
        super().__init__()
        if distance_type != "l1":
            raise ValueError(distance_type)

    def forward(self, labels):
        return (labels[:, None, :] - labels[None, :, :]).abs().sum(dim=-1)


class FeatureSimilarity(nn.Module):
    def __init__(self, similarity_type="l2"):
        super().__init__()
        if similarity_type != "l2":
            raise ValueError(similarity_type)

    def forward(self, features):
        return -(features[:, None, :] - features[None, :, :]).norm(2, dim=-1)


class RnCLoss(nn.Module):
    """Rank-N-Contrast.  Vectorised over the reference's ``for k in range(n-1)`` loop (:637-642)."""

    def __init__(self, temperature=2, label_diff="l1", feature_sim="l2"):
        super().__init__()
        self.t = temperature
        self.label_diff_fn, self.feature_sim_fn = LabelDifference(label_diff), FeatureSimilarity(feature_sim)

    def forward(self, features, labels):
        if len(features.shape) == 2 * len(labels.shape):
            features = torch.cat([features[:, 0], features[:, 1]], dim=0)
            labels = labels.repeat(2, 1)
        n = features.shape[0]
        if n < 2:
            return 0.0  # the reference's loop body never runs (:636-637)
        d = self.label_diff_fn(labels)
        logits = self.feature_sim_fn(features) / self.t
        logits = logits - logits.max(dim=1, keepdim=True).values.detach()
        off = ~torch.eye(n, dtype=torch.bool, device=logits.device)
        logits, d = logits[off].view(n, n - 1), d[off].view(n, n - 1)
        keep = (d[:, None, :] >= d[:, :, None]).to(logits.dtype)                 # [i, k, j]: d_ij >= d_ik
        denom = (keep * logits.exp()[:, None, :]).sum(dim=-1)                    # [i, k]
        return -((logits - denom.log()) / (n * (n - 1))).sum()


class GenerativeContrastiveLoss(nn.Module):
    def __init__(self, ds_contra_loss, gen_loss, pred_space_contra_loss, regulatory_weight, ds_regulatory_weight):
        super().__init__()
        self.ds_contra_loss, self.gen_loss, self.pred_space_contra_loss = ds_contra_loss, gen_loss, pred_space_contra_loss
        self.reg_weight, self.ds_reg_weight, self.gen_weight = regulatory_weight, ds_regulatory_weight, 1.0

    def forward(self, prediction, target, roi, final_representations, intermediate_extractions):
        gen = self.gen_loss(prediction, target, roi)
        reduced = gen.sum() if self.gen_loss.batch_reduction is None else gen
        ps = self.reg_weight * self.pred_space_contra_loss(*final_representations)
        ds = self.ds_reg_weight * self.ds_contra_loss(*intermediate_extractions)
        if ps.device != gen.device:
            ps = ps.to(gen.device)
        return self.gen_weight * reduced + ps + ds, gen, ps, ds
