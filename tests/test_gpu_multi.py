"""GPU, 2 ranks over NCCL (skipped with fewer than two devices): the REAL model under DataParallelEngine against one process
that sees the global batch (reference semantics: the loss SUMS over the batch, criterions.py:560, and
RnC ranks every sample against every other, :623-642).

Single-process equivalent of "BatchNorm statistics per rank" (what nn.DataParallel does): the two halves of the global batch
go through the model in two separate forward calls (each normalises over its own samples), the RnC features are concatenated,
one backward.  Compared: every all-reduced gradient, the set of parameters whose grad stays None, the order in which the ranks
issued their collectives.  Rank 0's samples select the positive prompt and rank 1's the negative one.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

pytestmark = pytest.mark.gpu

CASE = {"channels": [8, 16, 32, 64, 128], "shape": (32, 32, 32), "batch": 4, "seed": 41}


def _build(device):
    import coma_unet_b200 as cu
    from tests.golden import common
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, CASE["channels"], [2] * 5, latent_spaces=[2048] * 5, conditional=True,
                                       prompt_shape=CASE["shape"], compute_dtype=torch.float32)
    m.set_save_attn(None)
    return common.fill_deterministic(m, CASE["seed"]).to(device)


def _criterion():
    import coma_unet_b200 as cu
    from tests.golden import common
    gen = cu.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
    crit = cu.GenerativeContrastiveLoss(cu.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)
    crit.gen_loss.batch_reduction = None
    return crit


def _data():
    from tests.golden import common
    mri, tau, roi, covars, dicts = common.synthetic_batch(CASE["batch"], CASE["shape"], CASE["seed"])
    covars[:2, 0, 0], covars[2:, 0, 0] = 1.0, 0.0     # rank 0: positive prompt only; rank 1: negative prompt only
    return mri, tau, roi, covars, dicts


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from coma_unet_b200.parallel import DataParallelEngine
    torch.manual_seed(100 + rank)
    m = _build(dev)
    if rank == 1:                                      # replicas differ until the engine broadcasts rank 0's parameters
        with torch.no_grad():
            m.general_dynamic_prompt.add_(1.0)
    eng = DataParallelEngine(m, bucket_mb=0.05)        # small buckets: many collectives, so the order matters
    crit = _criterion()
    mri, tau, roi, covars, dicts = _data()
    sl = slice(rank * 2, rank * 2 + 2)
    logs = []
    for step in range(2):                              # step 2 runs on the rebuilt (learned) buckets
        for p in m.parameters():
            p.grad = None
        m.train(True)
        pred, proj, final = m(mri[sl].to(dev), covars[sl], roi_pred_dicts=dicts[sl], sample_roi_mask=roi[sl].to(dev))
        feats, labels = eng.gather_rnc(proj[-1], covars[sl, -1].float().to(dev))
        z = torch.zeros(final.size(), device=dev)
        loss, _, _, _ = crit(pred, tau[sl].to(dev), roi[sl].to(dev), (final, z, z), (feats, labels))
        loss.backward()
        eng.finish()
        torch.cuda.synchronize()
        logs.append([tuple(eng.buckets[b]) for b in eng.launch_log])     # no optimizer step: step 2 repeats step 1's gradients
    grads = {k: (None if p.grad is None else p.grad.detach().float().cpu()) for k, p in m.named_parameters()}
    out.put((rank, grads, logs, eng.launched_in_backward, len(eng.buckets)))
    dist.barrier()
    dist.destroy_process_group()


def _single_process():
    dev = torch.device("cuda", 0)
    m = _build(dev)
    crit = _criterion()
    mri, tau, roi, covars, dicts = _data()
    m.train(True)
    preds, feats, finals = [], [], []
    for sl in (slice(0, 2), slice(2, 4)):              # per-half BatchNorm statistics, one autograd graph
        pred, proj, final = m(mri[sl].to(dev), covars[sl], roi_pred_dicts=dicts[sl], sample_roi_mask=roi[sl].to(dev))
        preds.append(pred); feats.append(proj[-1]); finals.append(final)
    pred, feat, final = torch.cat(preds), torch.cat(feats), torch.cat(finals)
    z = torch.zeros(final.size(), device=dev)
    loss, _, _, _ = crit(pred, tau.to(dev), roi.to(dev), (final, z, z), (feat, covars[:, -1].float().to(dev)))
    loss.backward()
    return {k: (None if p.grad is None else p.grad.detach().float().cpu()) for k, p in m.named_parameters()}


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_two_nccl_ranks_match_the_global_batch():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, grads, logs, in_bwd, nb = out.get()
        got[r] = (grads, logs, in_bwd, nb)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = _single_process()
    assert got[0][1] == got[1][1], "ranks issued their all-reduces in different orders"
    assert got[0][2] > 0, "no bucket was launched while backward was still running (no overlap)"
    for rank in (0, 1):
        grads = got[rank][0]
        assert {k for k, g in grads.items() if g is None} == {k for k, g in want.items() if g is None}, rank
        worst = {}
        for k, g in want.items():
            if g is None:
                continue
            scale = float(g.abs().max())
            if scale < 1e-10:
                assert float(grads[k].abs().max()) < 1e-6, k
                continue
            worst[k] = float((grads[k] - g).abs().max()) / scale
        bad = {k: v for k, v in worst.items() if v > 5e-4}
        assert not bad, (rank, sorted(bad.items(), key=lambda kv: -kv[1])[:8])
    assert want["pos_dynamic_prompt"] is not None and want["neg_dynamic_prompt"] is not None
