"""One-process-per-GPU data parallelism for the training step.

Replaces the reference's (imported but disabled) ``torch.nn.DataParallel`` (attn_unet_data_parallel.py:32;
wrap sites commented out at validation.py:268-269) with what it would have meant, done the B200 way:

* every rank owns whole volumes, a full parameter replica and its own ``roi_pred_dicts`` sub-list;
  BatchNorm statistics stay per-rank (``nn.DataParallel`` semantics; ``SyncBatchNorm`` is unused in the reference);
  parameters and buffers are broadcast from rank 0 when the engine is built, so replicas start identical whatever
  each process seeded;
* gradients are **summed** (the reference loss sums over the batch, criterions.py:560) with bucketed NCCL
  all-reduces launched from post-accumulate-grad hooks while backward is still running (NVLink/NVSwitch, ~25 MB
  buckets in reverse parameter order, flat buffers the ``.grad`` tensors alias -- no pack/unpack copies);
* **the all-reduces are issued in the same order on every rank.**  Which hooks fire depends on the rank's own data
  (``pos_dynamic_prompt`` / ``neg_dynamic_prompt`` get a gradient only if a local sample selects them,
  attn_unet_data_parallel.py:638-639), so "launch a bucket when its last hook fires" would pair different buffers
  across ranks.  Parameters are therefore split into *early* ones, whose hook fires on every rank in every step,
  and *late* ones (declared by the model through ``data_dependent_parameters()`` / ``end_of_backward_parameters()``, plus every parameter that fired
  on no rank in the first step: ``reweigh*``, ``modulator*``, projection heads 0-3).  Early buckets are launched
  by a cursor, bucket b only after buckets 0..b-1; late buckets are launched in ``finish()``, in index order, and
  only if some rank used one of their parameters (a decision taken from the globally reduced mask, hence identical
  everywhere).  A parameter that misbehaves (early, but silent on some rank) only delays the cursor until
  ``finish()``; the order stays the same on every rank;
* parameters that received no gradient on ANY rank keep ``grad = None`` (AdamW then skips them exactly as in the
  single-process reference; SURVEY hard part 5) -- decided by one tiny MAX all-reduce of a used-bitmask, host to
  host (gloo), so the step never waits for the GPU;
* ``gather_rnc`` all-gathers the ``[B_local,512]`` features / ``[B_local,6]`` labels so ``RnCLoss`` ranks the global
  batch (criterions.py:623-642), with the matching backward.

Inference needs none of this: batches are split across ranks and nothing is exchanged.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

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
    def __init__(self, model: torch.nn.Module, world_size: int | None = None, bucket_mb: float = 25.0, group=None,
                 late: Optional[Iterable[torch.nn.Parameter]] = None, broadcast: bool = True, overlap: Optional[bool] = None):
        self.model, self.group = model, group
        # overlap=False: every bucket is reduced in finish(), after backward (A/B switch; env COMA_DP_OVERLAP=0)
        self.overlap = (os.environ.get("COMA_DP_OVERLAP", "1") != "0") if overlap is None else bool(overlap)
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        self.enabled = self.world > 1
        self.comm_bytes = 0          # gradient bytes all-reduced in the last step (bench.py reports it)
        self.timing = False          # bench.py: CUDA events around the wait for the all-reduce tail (exposed communication)
        self.timing_events = []
        self.launched_in_backward = 0
        if not self.enabled:
            return
        if broadcast:                # replicas start from rank 0's state, as DistributedDataParallel does
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.cap = int(float(os.environ.get("COMA_DP_BUCKET_MB", bucket_mb)) * 1024 * 1024)
        declared = late
        if declared is None:
            declared = []
            for name in ("data_dependent_parameters", "end_of_backward_parameters"):
                fn = getattr(model, name, None)
                declared += list(fn()) if callable(fn) else []
        ids = {id(p) for p in declared}
        self.late = torch.tensor([id(p) in ids for p in self.params], dtype=torch.bool)
        self.learned = False         # the never-firing parameters join ``late`` after the first step
        self._build_buckets()
        self.fired = torch.zeros(len(self.params), dtype=torch.uint8)      # host-side, filled by hooks
        # The used-parameter mask lives on the HOST (the hooks run when autograd executes a node, ahead of the GPU), so its MAX
        # all-reduce goes over a gloo group: an NCCL all-reduce would force a device->host read at the end of backward, i.e. the
        # host would wait for the whole step and the GPU would idle while the optimizer step and the next forward are enqueued.
        self.host_group = None
        if dist.is_initialized() and dist.get_backend(group) != "gloo":
            ranks = dist.get_process_group_ranks(group) if group is not None else None
            self.host_group = dist.new_group(ranks=ranks, backend="gloo")
        self.handles = []
        self.launch_log: List[int] = []      # bucket indices in the order their all-reduce was issued (tests compare ranks)
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))
        model.register_forward_pre_hook(lambda m, a: self.attach())
        self.attached = False

    # -- buckets -----------------------------------------------------------------------------------
    def _build_buckets(self):
        """Early parameters in reverse registration order (~ the order gradients become ready), then the late ones."""
        def fill(indices):
            out, cur, cur_bytes = [], [], 0
            for idx in indices:
                p = self.params[idx]
                nbytes = p.numel() * p.element_size()
                if cur and (cur_bytes + nbytes > self.cap or p.dtype != self.params[cur[0]].dtype):
                    out.append(cur)
                    cur, cur_bytes = [], 0
                cur.append(idx)
                cur_bytes += nbytes
            if cur:
                out.append(cur)
            return out

        order = list(reversed(range(len(self.params))))
        early = fill([i for i in order if not self.late[i]])
        tail = fill([i for i in order if self.late[i]])
        self.buckets = early + tail
        self.n_early = len(early)
        self.bucket_of, self.flat, self.views = {}, [], {}
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
        self._reset_step()

    def _reset_step(self):
        self.pending = [len(idxs) for idxs in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.cursor = 0
        self.handles = []
        self.launch_log = []
        self.comm_bytes = 0

    # -- per step --------------------------------------------------------------------------------
    def attach(self):
        """Reset the per-step bookkeeping; called before each forward.  Gradients are NOT pre-pointed at the flat buckets: autograd
        would then accumulate into them with one ``add_`` kernel per parameter (~280 tiny launches, ~0.8 ms of a 30 ms step);
        instead it hands every parameter a fresh gradient tensor for free and ``_launch`` gathers a whole bucket with ONE
        multi-tensor copy right before its all-reduce."""
        if not self.enabled or self.attached or not torch.is_grad_enabled():
            return
        self.fired.zero_()
        self._reset_step()
        self.attached = True

    def _make_hook(self, i):
        def hook(param):
            if not self.attached:
                return
            if self.fired[i]:                             # a second accumulation into an already gathered parameter
                if self.launched[self.bucket_of[i]]:
                    raise RuntimeError("DataParallelEngine: a gradient arrived after its bucket was all-reduced")
                return
            self.fired[i] = 1
            b = self.bucket_of[i]
            self.pending[b] -= 1
            # early buckets go out strictly in index order: bucket b waits for 0..b-1 (same sequence on every rank)
            while self.overlap and self.cursor < self.n_early and self.pending[self.cursor] == 0:
                self._launch(self.cursor)
                self.cursor += 1
        return hook

    def _gather(self, b):
        """Local gradients of bucket b -> its flat buffer (one multi-tensor copy), zeros for parameters this rank did not use; the
        ``.grad`` attributes then alias the buffer, so the optimizer reads the reduced values without another copy."""
        idxs = self.buckets[b]
        have = [i for i in idxs if self.params[i].grad is not None and self.params[i].grad is not self.views[i]]
        if have:
            torch._foreach_copy_([self.views[i] for i in have], [self.params[i].grad for i in have])
        for i in idxs:
            if self.params[i].grad is None:
                self.views[i].zero_()
            self.params[i].grad = self.views[i]

    def _launch(self, b):
        if not self.launched[b]:
            self.launched[b] = True
            self._gather(b)
            self.launch_log.append(b)
            self.comm_bytes += self.flat[b].numel() * self.flat[b].element_size()
            self.handles.append(dist.all_reduce(self.flat[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """After ``loss.backward()``: flush what the cursor held back, resolve the used-bitmask, reduce the late buckets
        some rank used (in index order), wait."""
        if not self.enabled:
            return
        self.launched_in_backward = len(self.launch_log)
        for b in range(self.n_early):
            self._launch(b)
        self.cursor = self.n_early
        used = self.fired.clone()
        dist.all_reduce(used, op=dist.ReduceOp.MAX, group=self.host_group if self.host_group is not None else self.group)
        for b in range(self.n_early, len(self.buckets)):
            if any(used[i] for i in self.buckets[b]):     # identical on every rank: ``used`` is the global mask
                self._launch(b)
            else:
                for i in self.buckets[b]:                 # used nowhere: no collective, the gradient stays None
                    self.params[i].grad = None
        if self.timing:    # the compute stream idles between these two events exactly as long as the all-reduce tail is exposed
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for h in self.handles:
            h.wait()                                  # stream-level wait: the host does not block
        if self.timing:
            e1.record()
            self.timing_events.append((e0, e1))
        for i in (used == 0).nonzero().flatten().tolist():
            self.params[i].grad = None
        self.attached = False
        if not self.learned:
            # parameters no rank used in the first step (never part of the graph: reweigh*, modulator*, projection heads 0-3) stop
            # holding the cursor back.  The mask is global, so every rank rebuilds the same buckets.
            self.learned = True
            never = (used == 0) & ~self.late
            if bool(never.any()):
                grads = {i: p.grad for i, p in enumerate(self.params)}
                self.late = self.late | never
                self._build_buckets()
                for i, g in grads.items():            # keep this step's reduced gradients valid for optimizer.step()
                    if g is not None:
                        self.views[i].copy_(g)
                        self.params[i].grad = self.views[i]

    def gather_rnc(self, features, labels):
        if not self.enabled:
            return features, labels
        return _AllGatherCat.apply(features, self.group), _AllGatherCat.apply(labels.contiguous(), self.group).detach()

    def all_reduce_scalars(self, *values: float) -> List[float]:
        """SUM over ranks of a few host scalars (epoch loss sums, sample counts): keeps rank-local decisions such as
        ReduceLROnPlateau identical on every replica."""
        if not self.enabled:
            return list(values)
        t = torch.tensor(values, dtype=torch.float64)
        if self.host_group is not None or dist.get_backend(self.group) == "gloo":
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.host_group if self.host_group is not None else self.group)
        else:
            d = t.to(self.params[0].device)
            dist.all_reduce(d, op=dist.ReduceOp.SUM, group=self.group)
            t = d.cpu()
        return t.tolist()

    def shard(self, items, rank: int):
        """Split a global batch (list or tensor, batch-first) into this rank's contiguous slice (inference / loaders)."""
        n = len(items)
        per = (n + self.world - 1) // self.world
        return items[rank * per:min(n, (rank + 1) * per)]
