"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's input preparation, the step right
before the hot path.

* ``resize_volume``   follows ``VolumeDataset.resize_volume`` (VolumeDataset.py:236-259): SimpleITK ``ResampleImageFilter`` with
  the input's origin and direction, identity transform, nearest-neighbour interpolation, output size
  ``round(size * spacing / new_spacing)`` and default pixel value ``volume.GetPixelIDValue()`` (the pixel-type enum, 8 for
  float32 images).  **parity unpinned**: SimpleITK is not installed here and the reference holds no fixture for it; the
  arithmetic is ITK's published one (output index -> physical point -> continuous input index in double, inside iff
  ``-0.5 <= c < size - 0.5`` per axis, nearest index ``floor(c + 0.5)``).
* ``pad_volume``      follows ``data_util.pad_volume`` (data_util.py:814-828).  **pinned**: tests/golden/prepare_golden.npz holds
  outputs of the reference's own function (tests/golden/make_prepare_golden.py extracts it from data_util.py and runs it).
* ``load_volume``     follows ``VolumeDataset.load_volume_file`` / ``apply_transforms`` (VolumeDataset.py:214-264) after the read.
* ``prepare_sample``  adds ``mri[roi == 0] = 0`` (VolumeDataset_ADNI_A4_combined.py:63-68).

Arrays are [z, y, x] (``sitk.GetArrayFromImage``); spacings are (x, y, z) as SimpleITK reports them.
"""
from __future__ import annotations

import numpy as np
import torch


def resize_volume(array, spacing, new_spacing=(2.0, 2.0, 2.0), default_value=8.0):
    """array [z, y, x] -> nearest-neighbour resample onto ``new_spacing`` (VolumeDataset.py:236-259)."""
    array = np.asarray(array)
    size_xyz = array.shape[::-1]
    out_size = [int(np.round(size_xyz[j] * (spacing[j] / new_spacing[j]))) for j in range(3)]          # :241-245
    index, inside = [], []
    for j in range(3):
        c = np.arange(out_size[j], dtype=np.float64) * (float(new_spacing[j]) / float(spacing[j]))
        inside.append((c >= -0.5) & (c < size_xyz[j] - 0.5))
        index.append(np.clip(np.floor(c + 0.5).astype(np.int64), 0, size_xyz[j] - 1))
    out = array[np.ix_(index[2], index[1], index[0])].copy()
    ok = inside[2][:, None, None] & inside[1][None, :, None] & inside[0][None, None, :]
    out[~ok] = default_value                                                                          # :252
    return out


def pad_volume(target_size):
    """data_util.py:814-828 (centred zero padding up to ``target_size``; only dim -2 is ever cropped)."""
    def pad(volume):
        n = volume.dim()
        dims = []
        for i in range(1, 4):
            before = max(0, (target_size[i - 1] - volume.size(dim=n - i)) // 2)
            after = max(0, target_size[i - 1] - volume.size(dim=n - i) - before)
            dims.extend([before, after])
        padded = torch.nn.functional.pad(volume, tuple(dims), "constant", 0)
        if padded.size(dim=-2) != target_size[1]:
            padded = padded[:, :, :target_size[1], :]
        return padded
    return pad


def load_volume(array, spacing, resize=True, pad_dims=(128, 128, 128), default_value=8.0):
    """VolumeDataset.py:214-233 after the file read: resample, to float32 [1, z, y, x], nan_to_num, pad."""
    if resize:
        array = resize_volume(array, spacing, (2.0, 2.0, 2.0), default_value)
    t = torch.from_numpy(np.ascontiguousarray(array)).to(dtype=torch.float32).unsqueeze(dim=0)
    t = torch.nan_to_num(t)
    if pad_dims is not None and t.size(dim=-3) != pad_dims[-3]:                                        # apply_transforms :261-264
        t = pad_volume(pad_dims)(t)
    return t


def prepare_sample(mri, tau, roi, spacing, resize=True, pad_dims=(128, 128, 128), default_value=8.0):
    """VolumeDataset_ADNI_A4_combined.py:63-68: the three volumes of one sample, MRI masked by the ROI map."""
    mri_t = load_volume(mri, spacing, resize, pad_dims, default_value)
    tau_t = load_volume(tau, spacing, resize, pad_dims, default_value)
    roi_t = load_volume(roi, spacing, resize, pad_dims, default_value)
    mri_t[roi_t == 0] = 0
    return mri_t, tau_t, roi_t
