"""Synthetic stand-in for the reference datasets with the same ``__getitem__`` contract:

    (mri[1,D,H,W] f32, tau[1,D,H,W] f32, roi[1,D,H,W] f32 FreeSurfer labels, (abeta, covars[1,6]), path)

(VolumeDataset_ADNI_A4_combined.py:86-91, VolumeDataset_Inference.py:145, VolumeDataset.py:382,646-647).
Covariate order ``[Abeta, Age, Sex, Education, Cognition, Tau_Meta]``; ``covars`` is float64 for the
"combined"/"inference" flavours and float32 for the "adni" flavour (VolumeDataset.py:427).  Value
distributions follow SURVEY.md section 8(d).  There is no network and no lab data here, so this is
what tests and bench.py feed the model.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch.utils.data import Dataset

from . import _lib
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


class DevicePrefetcher:
    """Device-side half of the reference's ``DataLoader(pin_memory=True)`` + ``.cuda(non_blocking=True)`` idiom
    (attn_unet_data_parallel.py:795-812): iterates batches (tuples whose tensors live in pinned host memory), issues each
    batch's host->device copies on a dedicated copy stream ``depth - 1`` batches ahead of the consumer, and hands out device
    tensors that the consumer's current stream may use at once.  The device buffers are a fixed ring of ``depth`` slots
    (allocated on first use, re-allocated if a batch changes shape), so the steady state never touches the allocator; a
    slot is refilled only after the work the consumer enqueued on it has finished.  Non-tensor items pass through.
    """

    def __init__(self, batches, device, depth=2):
        self.batches, self.device, self.depth = batches, torch.device(device), max(1, depth)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None] * self.depth          # per slot: [device tensors..., consumer-done event]

    def _issue(self, batch, slot):
        cur = self.slots[slot]
        same = cur is not None and len(cur[0]) == len(batch) and all(
            (not torch.is_tensor(t)) or (torch.is_tensor(d) and d.shape == t.shape and d.dtype == t.dtype)
            for t, d in zip(batch, cur[0]))
        if not same:
            bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) if torch.is_tensor(t) else None for t in batch]
            cur = self.slots[slot] = [bufs, None]
        if cur[1] is not None:
            self.stream.wait_event(cur[1])          # the consumer's work on this slot's previous batch
        with torch.cuda.stream(self.stream):
            for t, d in zip(batch, cur[0]):
                if torch.is_tensor(t):
                    d.copy_(t, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return slot, ready, tuple(d if torch.is_tensor(t) else t for t, d in zip(batch, cur[0]))

    def __iter__(self):
        from collections import deque
        inflight = deque()
        n = 0

        def hand_out():
            slot, ready, dev = inflight.popleft()
            torch.cuda.current_stream(self.device).wait_event(ready)
            return slot, dev

        def release(slot):
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self.slots[slot][1] = done

        for batch in self.batches:
            if len(inflight) >= self.depth:      # every slot holds a batch: hand the oldest to the consumer first
                slot, dev = hand_out()
                yield dev
                release(slot)
            inflight.append(self._issue(batch, n % self.depth))
            n += 1
        while inflight:
            slot, dev = hand_out()
            yield dev
            release(slot)


class HostSink:
    """Device->host read-back of per-step results into a small ring of pinned buffers on its own copy stream, so the
    read of step i overlaps the compute of step i+1 (the reference reads ``pred.cpu()`` synchronously,
    attn_unet_data_parallel.py:1226-1240).  ``put`` returns the pinned buffer that will hold the result once
    ``wait()`` (or the buffer's next reuse) has returned."""

    def __init__(self, shape, dtype=torch.float32, device="cuda", depth=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.bufs = [torch.empty(tuple(shape), dtype=dtype).pin_memory() for _ in range(max(1, depth))]
        self.done = [None] * len(self.bufs)
        self.i = 0

    def put(self, t):
        i = self.i
        self.i = (i + 1) % len(self.bufs)
        if self.done[i] is not None:
            self.done[i].synchronize()              # the buffer's previous read-back (depth steps old) has landed
        produced = torch.cuda.Event()
        produced.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(produced)
        with torch.cuda.stream(self.stream):
            self.bufs[i].copy_(t.detach().reshape(self.bufs[i].shape), non_blocking=True)
            t.record_stream(self.stream)
            self.done[i] = torch.cuda.Event()
            self.done[i].record(self.stream)
        return self.bufs[i]

    def wait(self):
        for ev in self.done:
            if ev is not None:
                ev.synchronize()


def prepare_geometry(in_shape, spacing, resize=True, new_spacing=(2.0, 2.0, 2.0), pad_dims=(128, 128, 128)):
    """Index geometry of the reference's per-sample preparation, per array axis (z, y, x) -- host integers only.

    ``in_shape`` is the [z, y, x] array shape, ``spacing``/``new_spacing`` are (x, y, z) as SimpleITK reports them.  Follows
    ``resize_volume`` (VolumeDataset.py:241-245: ``int(np.round(size * spacing / new_spacing))``), ``apply_transforms``
    (:261-264: pad only when the z extent differs from ``pad_dims[-3]``) and ``data_util.pad_volume`` (data_util.py:814-828:
    ``target_size[i]`` belongs to array axis ``-1 - i``, centred padding, only the y axis is ever cropped, at its end).
    Returns ``(res_size, out_size, pad_before, ratio)``.
    """
    res, ratio = [], []
    for ax in range(3):
        j = 2 - ax
        if resize:
            res.append(int(np.round(in_shape[ax] * (spacing[j] / new_spacing[j]))))
            ratio.append(float(new_spacing[j]) / float(spacing[j]))
        else:
            res.append(int(in_shape[ax]))
            ratio.append(1.0)
    out, before = list(res), [0, 0, 0]
    if pad_dims is not None and res[0] != pad_dims[-3]:
        for ax in range(3):
            target = pad_dims[2 - ax]
            before[ax] = max(0, (target - res[ax]) // 2)
            out[ax] = res[ax] + before[ax] + max(0, target - res[ax] - before[ax])
        if out[1] != pad_dims[1]:
            out[1] = min(out[1], pad_dims[1])
    return res, out, before, ratio


def prepare_volumes(mri=None, tau=None, roi=None, spacing=(1.0, 1.0, 1.0), resize=True, pad_dims=(128, 128, 128),
                    default_value=8.0):
    """GPU replacement for what the reference datasets do to one sample between the file read and the model (SURVEY 8f
    rank 4): ``load_volume_file`` for each of the three volumes (2 mm nearest-neighbour resample, ``nan_to_num``, centred zero
    padding; VolumeDataset.py:214-264, data_util.py:814-828) and ``mri[roi == 0] = 0``
    (VolumeDataset_ADNI_A4_combined.py:63-68), in ONE kernel launch over the output voxels (``coma_prepare_volumes``).

    Inputs are fp32 CUDA arrays [z, y, x] as ``sitk.GetArrayFromImage`` lays them out (any of them may be None); ``default_value``
    is what the reference passes as the out-of-image value, ``volume.GetPixelIDValue()`` (8 for float32 images).
    Returns ``(mri, tau, roi)`` as [1, Z, Y, X] fp32 tensors (None where the input was None).
    """
    given = [t for t in (mri, tau, roi) if t is not None]
    if not given:
        raise ValueError("prepare_volumes: no volume given")
    shape = tuple(given[0].shape[-3:])
    for t in given:
        if not t.is_cuda:
            raise RuntimeError("prepare_volumes runs on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.float32 or tuple(t.shape[-3:]) != shape or t.numel() != shape[0] * shape[1] * shape[2]:
            raise ValueError("prepare_volumes: volumes must be fp32 and share one [z, y, x] shape")
    res, out, before, ratio = prepare_geometry(shape, spacing, resize, (2.0, 2.0, 2.0), pad_dims)
    if min(res) <= 0:
        raise ValueError(f"prepare_volumes: empty resampled volume {res}")
    srcs = [None if t is None else t.contiguous() for t in (mri, tau, roi)]
    outs = [None if t is None else torch.empty((1, *out), dtype=torch.float32, device=t.device) for t in (mri, tau, roi)]
    i3 = C.c_int32 * 3
    a = _lib.PrepareArgs(*[_lib.ptr(t) for t in srcs], *[_lib.ptr(t) for t in outs], i3(*shape), i3(*res), i3(*out),
                         i3(*before), (C.c_double * 3)(*ratio), float(default_value))
    with torch.cuda.device(given[0].device):
        _lib.call("coma_prepare_volumes", C.byref(a), _lib.stream())
    return tuple(outs)
