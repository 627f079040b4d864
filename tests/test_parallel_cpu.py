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


def make_data():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 6, generator=g)
    y = torch.randn(4, 1, generator=g)
    labels = torch.rand(4, 6, generator=g)
    flag = torch.tensor([1, 1, 1, 1])    # no sample selects `neg` on any rank -> its grad must stay None
    return x, y, labels, flag


def loss_fn(pred, y, feats, labels):
    return ((pred - y) ** 2).sum() + RnCLoss()(feats, labels)      # SUM over the batch + batch-coupled term


def single_process():
    torch.manual_seed(1)
    m = Toy()
    x, y, labels, flag = make_data()
    pred, f = m(x, flag)
    loss_fn(pred, y, f, labels).backward()
    return {k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()}


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1)
    m = Toy()
    eng = DataParallelEngine(m, world_size=world, bucket_mb=0.0005)      # tiny buckets -> several of them
    assert len(eng.buckets) > 2
    x, y, labels, flag = make_data()
    sl = slice(rank * 2, rank * 2 + 2)
    for step in range(2):                                                 # second step checks re-attachment
        for p in m.parameters():
            p.grad = None
        pred, f = m(x[sl], flag[sl])
        feats, lab = eng.gather_rnc(f, labels[sl])
        loss_fn(pred, y[sl], feats, lab).backward()
        eng.finish()
    if rank == 0:
        out.put({k: (None if p.grad is None else p.grad.tolist()) for k, p in m.named_parameters()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_match_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = single_process()
    assert got["unused"] is None and want["unused"] is None
    assert got["neg"] is None and want["neg"] is None
    for k, g in want.items():
        if g is not None:
            assert torch.allclose(torch.tensor(got[k]), g, rtol=1e-5, atol=1e-6), k


def test_single_rank_engine_is_a_noop():
    m = Toy()
    eng = DataParallelEngine(m, world_size=1)
    f = torch.randn(2, 8)
    assert eng.gather_rnc(f, f)[0] is f
    eng.finish()
    assert eng.shard(list(range(8)), 0) == list(range(8))
