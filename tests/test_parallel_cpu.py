"""CPU, gloo, world_size 2: the data-parallel engine's host logic (bucketing, overlap hooks, SUM reduction,
used-parameter bitmask, RnC all-gather with its backward) against a single-process run on the concatenated batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from coma_unet_b200.criterions import RnCLoss
from coma_unet_b200.parallel import DataParallelEngine


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = nn.Sequential(nn.Linear(6, 16), nn.ReLU(), nn.Linear(16, 8))
        self.head = nn.Linear(8, 1)
        self.pos = nn.Parameter(torch.ones(8))      # used only by samples with flag == 1 (like pos_dynamic_prompt)
        self.neg = nn.Parameter(torch.ones(8))
        self.unused = nn.Parameter(torch.ones(3))   # never used (like `reweigh`)

    def forward(self, x, flag):
        f = self.body(x)
        prompt = torch.stack([self.pos if fl == 1 else self.neg for fl in flag.tolist()])
        return self.head(f * prompt), f

    def data_dependent_parameters(self):
        return [self.pos, self.neg]

    def end_of_backward_parameters(self):
        """Like the fused FiLM MLPs of the real model: parameters the engine should not wait for in the early buckets."""
        return list(self.head.parameters()) if getattr(self, "declare_head_late", False) else []


def make_data(flags=(1, 1, 1, 1)):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 6, generator=g)
    y = torch.randn(4, 1, generator=g)
    labels = torch.rand(4, 6, generator=g)
    flag = torch.tensor(flags)    # (1,1,1,1): no sample selects `neg` on any rank -> its grad must stay None
    return x, y, labels, flag


def loss_fn(pred, y, feats, labels):
    return ((pred - y) ** 2).sum() + RnCLoss()(feats, labels)      # SUM over the batch + batch-coupled term


def single_process(flags=(1, 1, 1, 1)):
    torch.manual_seed(1)
    m = Toy()
    x, y, labels, flag = make_data(flags)
    pred, f = m(x, flag)
    loss_fn(pred, y, f, labels).backward()
    return {k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()}


def worker(rank, world, port, out, flags, bucket_mb, declare, head_late=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1 + 7 * rank)                                       # replicas differ until the engine broadcasts rank 0's
    m = Toy()
    m.declare_head_late = head_late
    eng = DataParallelEngine(m, world_size=world, bucket_mb=bucket_mb, late=None if declare else [])
    assert len(eng.buckets) > 2
    if head_late:                                                         # the declared parameters sit in late buckets only
        late_ids = {id(p) for b in range(eng.n_early, len(eng.buckets)) for p in (eng.params[i] for i in eng.buckets[b])}
        assert all(id(p) in late_ids for p in m.head.parameters())
    x, y, labels, flag = make_data(flags)
    sl = slice(rank * 2, rank * 2 + 2)
    logs = []
    for step in range(3):                                                 # later steps check re-attachment / rebuilt buckets
        for p in m.parameters():
            p.grad = None
        pred, f = m(x[sl], flag[sl])
        feats, lab = eng.gather_rnc(f, labels[sl])
        loss_fn(pred, y[sl], feats, lab).backward()
        eng.finish()
        logs.append([tuple(eng.buckets[b]) for b in eng.launch_log])      # the parameter sets reduced, in issue order
    out.put((rank, {k: (None if p.grad is None else p.grad.tolist()) for k, p in m.named_parameters()}, logs,
             eng.all_reduce_scalars(float(rank + 1), 2.0)))
    dist.barrier()
    dist.destroy_process_group()


def run_two_ranks(flags, bucket_mb=0.0005, declare=True, head_late=False):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, out, flags, bucket_mb, declare, head_late)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r, (g, logs, sc)) for r, g, logs, sc in (out.get(), out.get()))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return got


def check_against_single_process(got, flags):
    want = single_process(flags)
    for rank in (0, 1):
        grads = got[rank][0]
        for k, g in want.items():
            if g is None:
                assert grads[k] is None, (rank, k)
            else:
                assert grads[k] is not None, (rank, k)
                assert torch.allclose(torch.tensor(grads[k]), g, rtol=1e-5, atol=1e-6), (rank, k)
    assert got[0][1] == got[1][1], "ranks issued their all-reduces in different orders"
    assert got[0][2] == got[1][2] == [3.0, 4.0]


def test_two_rank_gradients_match_single_process():
    flags = (1, 1, 1, 1)
    got = run_two_ranks(flags)
    assert got[0][0]["unused"] is None and got[0][0]["neg"] is None
    check_against_single_process(got, flags)


def test_ranks_that_select_different_prompts_reduce_in_the_same_order():
    """rank 0's samples select `pos`, rank 1's select `neg`; one parameter per bucket.  With launch-on-last-hook
    the `pos` bucket of rank 0 was paired with the `neg` bucket of rank 1."""
    flags = (1, 1, 0, 0)
    check_against_single_process(run_two_ranks(flags, bucket_mb=1e-6), flags)
    # an UNDECLARED data-dependent parameter only costs overlap (the cursor waits for finish()), never correctness
    check_against_single_process(run_two_ranks(flags, bucket_mb=1e-6, declare=False), flags)


def test_parameters_declared_end_of_backward_are_reduced_with_the_late_buckets():
    """model.end_of_backward_parameters() (the real model: the FiLM MLPs, whose gradients all come from one autograd node at the
    end of backward) joins the late set: same gradients, same launch order on both ranks."""
    flags = (1, 0, 0, 1)
    check_against_single_process(run_two_ranks(flags, bucket_mb=1e-6, head_late=True), flags)


def test_single_rank_engine_is_a_noop():
    m = Toy()
    eng = DataParallelEngine(m, world_size=1)
    f = torch.randn(2, 8)
    assert eng.gather_rnc(f, f)[0] is f
    eng.finish()
    assert eng.shard(list(range(8)), 0) == list(range(8))
