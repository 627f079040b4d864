"""CUDA-graph replay of the two hot steps: the training step (forward + loss + backward + gradient all-reduce + AdamW) and the
no-grad inference forward.

A batch-4 training step is ~900 kernel launches enqueued from Python (ctypes ABI calls, autograd nodes, ATen glue): 33.4 ms of
kernels in a 34.9 ms step, i.e. the host is barely ahead of the GPU and every allocator hiccup or interpreter pause shows up as an
idle GPU.  The step has static shapes and -- since round 1 -- no host synchronisation, so it is captured once per configuration
and replayed: no Python between the kernels, and the inter-kernel gaps go away.

What makes the captured step equivalent to the eager one (reference loop: attn_unet_data_parallel.py:812-885):

* inputs live in static device buffers that ``__call__`` refreshes before every replay (volumes by device copy, the covariates
  and the ROI look-up table from a pinned staging buffer);
* the only data-dependent HOST decision of the step -- which of ``pos_dynamic_prompt`` / ``neg_dynamic_prompt`` get a gradient
  (``grad is None`` for a prompt no sample selects, :638-639, which AdamW then skips) -- is made on the host before the replay
  and is the graph's key: one graph per (positive used, negative used) combination, captured lazily, sharing one memory pool.
  With several ranks the key is the GLOBAL usage (one tiny host all-reduce per step), so every rank replays the same
  sequence of collectives;
* the optimizer is fused AdamW with ``capturable=True`` and a tensor learning rate, so schedulers keep working
  (``ReduceLROnPlateau`` fills the tensor in place);
* NCCL all-reduces launched from the autograd hooks are captured with the step (``DataParallelEngine`` issues them in a fixed
  order); the engine's host-side mask exchange runs at capture time only -- its outcome is implied by the key.

Eager code that runs on the same model afterwards (validation, checkpoints) must call ``sync_eager()`` first: parameter
updates made inside a replay do not bump the tensors' version counters that the packed-weight caches are keyed on.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops


def _host_flags(covars) -> Tuple[bool, bool]:
    c0 = covars.reshape(covars.shape[0], -1)[:, 0]
    if c0.is_cuda:
        c0 = c0.cpu()                      # 4 floats; the inputs of a step do not depend on the previous step's kernels
    pos = (c0 == 1)
    return bool(pos.any()), bool((~pos).any())


class _StaticInputs:
    """Device buffers a graph reads + the pinned staging the small host-side inputs go through."""

    def __init__(self, model, mri, roi, tau, covars):
        dev = next(model.parameters()).device
        B = mri.shape[0]
        self.mri = torch.empty(mri.shape, dtype=torch.float32, device=dev)
        self.roi = torch.empty(roi.shape, dtype=torch.float32, device=dev)
        self.tau = None if tau is None else torch.empty(tau.shape, dtype=torch.float32, device=dev)
        self.n_cov = covars.reshape(B, -1).shape[1]
        n_roi = len(model.roi_indices)
        # one block: [B, n_cov] covariates followed by the [B, n_roi, 2] ROI table, uploaded through kernel parameters
        # (coma_upload_small): no copy engine, no pinned staging, nothing to race
        self.stage = torch.empty(B * self.n_cov + B * n_roi * 2, dtype=torch.float32)
        self.small = torch.empty(B * self.n_cov + B * n_roi * 2, dtype=torch.float32, device=dev)
        self.covars = self.small[:B * self.n_cov].view(B, 1, self.n_cov)
        self.lut = self.small[B * self.n_cov:].view(B, n_roi, 2)

    @staticmethod
    def _refresh(dst, src):
        # device -> device through an SM kernel, not cudaMemcpyAsync: a copy-engine D2D queues behind the (2.5 ms) H2D upload of
        # the NEXT batch that a prefetcher has in flight, which serialised upload and replay (measured: 12.1 -> 14.2 ms per step)
        if src.is_cuda and src.dtype == dst.dtype and src.shape == dst.shape:
            torch.mul(src, 1, out=dst)
        else:
            dst.copy_(src, non_blocking=True)

    def fill(self, model, mri, roi, tau, covars, roi_pred_dicts):
        self._refresh(self.mri, mri)
        self._refresh(self.roi, roi)
        if self.tau is not None:
            self._refresh(self.tau, tau)
        st = self.stage
        B = mri.shape[0]
        st[:B * self.n_cov].copy_(covars.reshape(-1).to(device="cpu", dtype=torch.float32))
        st[B * self.n_cov:].copy_(torch.from_numpy(model.roi_lut_host(roi_pred_dicts)).reshape(-1))
        _lib.call("coma_upload_small", self.small.data_ptr(), st.data_ptr(), st.numel(), _lib.stream())


class GraphedTrainStep:
    """``loss = step(mri, tau, roi, covars, roi_pred_dicts)``: one optimizer step, replayed from a CUDA graph.

    ``optimizer`` must be ``torch.optim.AdamW(..., fused=True, capturable=True)`` (``make_optimizer`` builds it); ``engine`` is
    the model's ``DataParallelEngine`` (world size 1 is fine).  Returns the step's loss as a device tensor; ``gen_loss`` holds
    the per-sample generative losses of the same step.
    """

    def __init__(self, model, criterion, optimizer, engine, warmup: int = 2):
        self.model, self.criterion, self.optimizer, self.engine = model, criterion, optimizer, engine
        self.warmup = max(int(warmup), 1)
        self.graphs: Dict[Tuple[bool, bool], torch.cuda.CUDAGraph] = {}
        self.outputs: Dict[Tuple[bool, bool], tuple] = {}
        self.launches: Dict[Tuple[bool, bool], int] = {}
        self.inputs: Optional[_StaticInputs] = None
        self.pool = None
        self.steps_seen = 0
        self.stream = torch.cuda.Stream(next(model.parameters()).device)
        self.gen_loss = None
        self.replays = 0

    @staticmethod
    def make_optimizer(model, lr: float):
        dev = next(model.parameters()).device
        return torch.optim.AdamW(model.parameters(), lr=torch.tensor(float(lr), device=dev), fused=True, capturable=True)

    # -- the step itself (eager warm-up, capture and ragged batches run the same code) ---------------------------------
    def _step(self, key, src=None):
        m = self.model
        mri, tau, roi, covars, lut = src if src is not None else (self.inputs.mri, self.inputs.tau, self.inputs.roi,
                                                                  self.inputs.covars, self.inputs.lut)
        self.optimizer.zero_grad(set_to_none=True)
        m._prompt_use_override = key
        try:
            pred, projected, final_repr = m(mri, covars, roi_pred_dicts=lut, sample_roi_mask=roi)[:3]
        finally:
            m._prompt_use_override = None
        cov_dev = covars.to(device=pred.device, dtype=torch.float32)
        feats, labels = self.engine.gather_rnc(projected[-1], cov_dev[:, -1])      # [B, 6] labels (:842-845)
        zeros = torch.zeros(final_repr.size(), device=pred.device)
        loss, gen, _, _ = self.criterion(pred, tau, roi, (final_repr, zeros, zeros), (feats, labels))
        loss.backward()
        self.engine.finish()
        self.optimizer.step()
        return loss.detach(), gen.detach()

    def eager(self, mri, tau, roi, covars, roi_pred_dicts):
        """One step WITHOUT replay, for a batch whose shape the graphs were not captured for (a ragged last batch).  It runs on the
        runner's stream: autograd binds every parameter's gradient accumulator to the stream it was first used on, and a capture
        that meets accumulators bound to another stream fails."""
        key = self._global_key(covars)
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            out = self._step(key, (mri, tau, roi, covars, roi_pred_dicts))
        cur.wait_stream(self.stream)
        self.gen_loss = out[1]
        return out[0]

    def _global_key(self, covars) -> Tuple[bool, bool]:
        pos, neg = _host_flags(covars)
        if self.engine.enabled:
            pos, neg = (v > 0 for v in self.engine.all_reduce_scalars(float(pos), float(neg)))
        return bool(pos), bool(neg)

    def matches(self, mri) -> bool:
        """Whether a batch has the shape the graphs were captured for (a ragged last batch has to run eagerly)."""
        return self.inputs is None or tuple(self.inputs.mri.shape) == tuple(mri.shape)

    def __call__(self, mri, tau, roi, covars, roi_pred_dicts):
        if self.inputs is None:
            self.inputs = _StaticInputs(self.model, mri, roi, tau, covars)
        if not self.matches(mri):
            raise ValueError(f"GraphedTrainStep was built for batches of shape {tuple(self.inputs.mri.shape)}, got {tuple(mri.shape)}")
        key = self._global_key(covars)
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.inputs.fill(self.model, mri, roi, tau, covars, roi_pred_dicts)
            if self.steps_seen < self.warmup:          # eager: sizes every cache / bucket and the optimizer state before the capture
                self.steps_seen += 1
                out = self._step(key)
            else:
                if key not in self.graphs:
                    self._capture(key)
                self.graphs[key].replay()
                self.replays += 1
                _lib.launches += self.launches[key]
                out = tuple(t.clone() for t in self.outputs[key])      # the graph's own output buffers are rewritten by the next replay
        cur.wait_stream(self.stream)
        self.model._coma_stale_caches = True
        self.gen_loss = out[1]
        return out[0]

    def _capture(self, key):
        ops.invalidate_weight_caches(self.model)      # the packing kernels must be part of the graph (weights change every replay)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = _lib.launches
        with torch.cuda.graph(g, pool=self.pool, stream=self.stream, capture_error_mode="thread_local"):
            out = self._step(key)
        if self.pool is None:
            self.pool = g.pool()
        self.launches[key] = _lib.launches - l0
        _lib.launches = l0
        self.graphs[key], self.outputs[key] = g, out
        # the capture itself executed nothing: the step it stands for runs with the first replay

    def sync_eager(self):
        """Call before running eager code (validation, another optimizer) on the model after replays."""
        torch.cuda.current_stream().wait_stream(self.stream)
        ops.invalidate_weight_caches(self.model)
        self.model._coma_stale_caches = False


class GraphedInference:
    """``pred = infer(mri, covars, roi_pred_dicts, roi)``: the no-grad forward replayed from a CUDA graph (eval mode, fixed shapes).
    The graph holds the packed weights of the moment it was captured: call ``reset()`` after the parameters change."""

    def reset(self):
        self.graph, self.out, self.seen = None, None, 0

    def __init__(self, model, warmup: int = 2):
        self.model, self.warmup = model, max(int(warmup), 1)
        self.graph, self.out, self.inputs, self.seen, self.launches = None, None, None, 0, 0
        self.stream = torch.cuda.Stream(next(model.parameters()).device)

    def _forward(self):
        inp = self.inputs
        with torch.no_grad():
            return self.model(inp.mri, inp.covars, roi_pred_dicts=inp.lut, sample_roi_mask=inp.roi)

    def __call__(self, mri, covars, roi_pred_dicts, roi):
        if self.model.training or self.model.embeddings_out:
            raise RuntimeError("GraphedInference replays the eval-mode forward that returns the prediction only")
        if self.inputs is None:
            self.inputs = _StaticInputs(self.model, mri, roi, None, covars)
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.inputs.fill(self.model, mri, roi, None, covars, roi_pred_dicts)
            if self.seen < self.warmup:
                self.seen += 1
                out = self._forward()
            else:
                if self.graph is None:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    l0 = _lib.launches
                    with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                        self.out = self._forward()
                    self.launches = _lib.launches - l0
                    _lib.launches = l0
                    self.graph = g
                self.graph.replay()
                _lib.launches += self.launches
                out = self.out.clone()       # 67 MB at batch 8: 0.02 ms; asynchronous consumers (HostSink) must not race the next replay
        cur.wait_stream(self.stream)
        return out
