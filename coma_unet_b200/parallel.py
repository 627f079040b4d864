"""One-process-per-GPU data parallelism for the training step.

Replaces the reference's (imported but disabled) ``torch.nn.DataParallel`` (attn_unet_data_parallel.py:32;
wrap sites commented out at validation.py:268-269) with what it would have meant, done the B200 way:

* every rank owns whole volumes, a full parameter replica and its own ``roi_pred_dicts`` sub-list;
  BatchNorm statistics stay per-rank (``nn.DataParallel`` semantics; ``SyncBatchNorm`` is unused in the reference);
* gradients are **summed** (the reference loss sums over the batch, criterions.py:560) with bucketed NCCL
  all-reduces launched from post-accumulate-grad hooks while backward is still running (NVLink/NVSwitch, ~25 MB
  buckets in reverse parameter order, flat buffers the ``.grad`` tensors alias -- no pack/unpack copies);
* parameters that received no gradient on ANY rank keep ``grad = None`` (AdamW then skips them exactly as in the
  single-process reference: ``reweigh*``, ``modulator*``, unused prompts, projection heads 0-3; SURVEY hard part 5)
  -- decided by one tiny MAX all-reduce of a used-bitmask, host to host (gloo), so the step never waits for the GPU;
* ``gather_rnc`` all-gathers the ``[B_local,512]`` features / ``[B_local,6]`` labels so ``RnCLoss`` ranks the global
  batch (criterions.py:623-642), with the matching backward.

Inference needs none of this: batches are split across ranks and nothing is exchanged.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class _AllGatherCat(torch.autograd.Function):
    """cat(all_gather(x)) whose backward returns d(sum over ranks of identical losses)/dx / world_size."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        ctx.group, ctx.rank, ctx.n = group, dist.get_rank(group), x.shape[0]
        return torch.cat(parts, dim=0)

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=ctx.group)
        world = dist.get_world_size(ctx.group)
        return grad[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n] / world, None


class DataParallelEngine:
    def __init__(self, model: torch.nn.Module, world_size: int | None = None, bucket_mb: float = 25.0, group=None):
        self.model, self.group = model, group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        self.enabled = self.world > 1
        if not self.enabled:
            return
        # buckets in reverse registration order (~ the order gradients become ready)
        cap = int(bucket_mb * 1024 * 1024)
        self.buckets, cur, cur_bytes = [], [], 0
        for idx in reversed(range(len(self.params))):
            p = self.params[idx]
            nbytes = p.numel() * p.element_size()
            if cur and (cur_bytes + nbytes > cap or p.dtype != self.params[cur[0]].dtype):
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(idx)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {}
        self.flat, self.views = [], {}
        for b, idxs in enumerate(self.buckets):
            p0 = self.params[idxs[0]]
            flat = torch.zeros(sum(self.params[i].numel() for i in idxs), device=p0.device, dtype=p0.dtype)
            off = 0
            for i in idxs:
                n = self.params[i].numel()
                self.views[i] = flat[off:off + n].view_as(self.params[i])
                self.bucket_of[i] = b
                off += n
            self.flat.append(flat)
        self.fired = torch.zeros(len(self.params), dtype=torch.uint8)      # host-side, filled by hooks
        # The used-parameter mask lives on the HOST (the hooks run when autograd executes a node, ahead of the GPU), so its MAX
        # all-reduce goes over a gloo group: an NCCL all-reduce would force a device->host read at the end of backward, i.e. the
        # host would wait for the whole step and the GPU would idle while the optimizer step and the next forward are enqueued.
        self.host_group = None
        if dist.is_initialized() and dist.get_backend(group) != "gloo":
            ranks = dist.get_process_group_ranks(group) if group is not None else None
            self.host_group = dist.new_group(ranks=ranks, backend="gloo")
        self.pending = [len(idxs) for idxs in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.handles = []
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))
        model.register_forward_pre_hook(lambda m, a: self.attach())
        self.attached = False

    # -- per step --------------------------------------------------------------------------------
    def attach(self):
        """Point every ``.grad`` at its slice of the (zeroed) flat bucket; called before each forward."""
        if not self.enabled or self.attached or not torch.is_grad_enabled():
            return
        for flat in self.flat:
            flat.zero_()
        for i, p in enumerate(self.params):
            p.grad = self.views[i]
        self.fired.zero_()
        self.pending = [len(idxs) for idxs in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.handles = []
        self.attached = True

    def _make_hook(self, i):
        def hook(param):
            if not self.attached:
                return
            if param.grad is not self.views[i]:           # autograd replaced the tensor: fold it back into the bucket
                self.views[i].copy_(param.grad)
                param.grad = self.views[i]
            if self.fired[i]:
                return
            self.fired[i] = 1
            b = self.bucket_of[i]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        if not self.launched[b]:
            self.launched[b] = True
            self.handles.append(dist.all_reduce(self.flat[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """After ``loss.backward()``: flush buckets with never-firing parameters, resolve the used-bitmask, wait."""
        if not self.enabled:
            return
        for b in range(len(self.buckets)):
            self._launch(b)
        used = self.fired.clone()
        dist.all_reduce(used, op=dist.ReduceOp.MAX, group=self.host_group if self.host_group is not None else self.group)
        for h in self.handles:
            h.wait()                                  # stream-level wait: the host does not block
        for i in (used == 0).nonzero().flatten().tolist():
            self.params[i].grad = None
        self.attached = False

    def gather_rnc(self, features, labels):
        if not self.enabled:
            return features, labels
        return _AllGatherCat.apply(features, self.group), _AllGatherCat.apply(labels.contiguous(), self.group).detach()

    def shard(self, items, rank: int):
        """Split a global batch (list or tensor, batch-first) into this rank's contiguous slice (inference / loaders)."""
        n = len(items)
        per = (n + self.world - 1) // self.world
        return items[rank * per:min(n, (rank + 1) * per)]
