"""Step loop equivalent to the reference's ``train_dp`` (attn_unet_data_parallel.py:696-1032), SURVEY.md section 8(f) rank 1.

Same call shape and the same state: ``criterion.gen_loss.batch_reduction = None`` (:717), AdamW(lr) +
ReduceLROnPlateau('min', patience=5) unless resuming (:729-737), one optimizer step per batch (:884-885),
``scheduler.step(epoch_loss / num_samples)`` per epoch (:921), the checkpoint dict
``{epoch, model_state_dict, optimizer_state_dict, loss, scheduler_state_dict}`` (``loss`` = the last batch's loss tensor,
as in the reference) written every epoch to ``<save_path>/checkpoints/checkpoint_latest_epoch.pth`` plus a numbered copy
whenever ``epoch % checkpoint_iter == 0`` (epochs 0, 5, 10, ...; :943-955), and a validation pass whenever
``epoch % val_iter == 0`` (:957-963).

What is deliberately different: no host synchronisation inside the step loop.  The reference calls ``loss.item()`` and
per-sample ``gen_loss[b].item()`` every batch (:892,901-910) and logs full-tensor norms (criterions.py:203-204); here the
running sums live on the device and are read once per epoch.  ROI prediction dicts come from ``roi_pred_fn(paths)``
(the reference reads lab-private JSON lookups, :708-710,809-810).  With ``torch.distributed`` initialised the step is
data-parallel through ``DataParallelEngine`` (gradient SUM all-reduce overlapped with backward, global-batch RnC): the
engine broadcasts rank 0's parameters and buffers when it is built, and the epoch loss sums / sample counts are summed over
ranks before ``scheduler.step`` so that ReduceLROnPlateau takes the same decision on every replica.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
from torch.optim.lr_scheduler import ReduceLROnPlateau

from .graph import GraphedTrainStep
from .parallel import DataParallelEngine


def _unpack(batch):
    """Reference loaders yield (anchor, positive, negative) triplets (:784); plain dataset tuples are accepted too."""
    if isinstance(batch, (tuple, list)) and len(batch) == 3 and isinstance(batch[0], (tuple, list)) and len(batch[0]) == 5:
        batch = batch[0]
    mri, tau, roi, (abeta, covars), paths = batch
    return mri, tau, roi, abeta, covars, list(paths)


def validate(model, loader, roi_pred_fn, device):
    """Mean absolute error over the loader (the full metric suite of contrastive_test is SURVEY 8(f) rank 3)."""
    was_training = model.training
    model.eval()
    model.set_training(False)
    total = torch.zeros((), device=device)
    count = 0
    with torch.no_grad():
        for batch in loader:
            mri, tau, roi, _, covars, paths = _unpack(batch)
            mri, tau, roi = mri.to(device, non_blocking=True), tau.to(device, non_blocking=True), roi.to(device, non_blocking=True)
            pred = model(mri, covars, roi_pred_dicts=roi_pred_fn(paths), sample_roi_mask=roi)
            total += (pred - tau).abs().mean(dim=(1, 2, 3, 4)).sum()
            count += mri.shape[0]
    model.train(was_training)
    model.set_training(was_training)
    return float(total / max(count, 1))


def train_dp(model, criterion, train_loader, validation_loader, epochs, lr, save_path="", cuda_id=0, pred_sample_file="",
             from_checkpoint=False, **kwargs):
    device = torch.device("cuda", cuda_id)
    roi_pred_fn = kwargs["roi_pred_fn"]
    val_iter = kwargs.get("val_iter", 5)
    checkpoint_iter = kwargs.get("checkpoint_iter", 5)
    cuda_graph = bool(kwargs.get("cuda_graph", False))       # replay the step from a CUDA graph (coma_unet_b200.graph)
    criterion.gen_loss.batch_reduction = None
    start_epoch = 0
    if from_checkpoint:
        optimizer = kwargs["optimizer"]
        start_epoch = kwargs["start_epoch"]
        scheduler = kwargs.get("scheduler") or ReduceLROnPlateau(optimizer, "min", patience=5, factor=0.2)
    else:
        if cuda_graph:      # capturable fused AdamW with a tensor learning rate (ReduceLROnPlateau fills it in place)
            optimizer = GraphedTrainStep.make_optimizer(model, lr)
        else:
            optimizer = torch.optim.AdamW(model.parameters(), lr, fused=next(model.parameters()).is_cuda)   # one multi-tensor kernel per step
        scheduler = ReduceLROnPlateau(optimizer, "min", patience=5)
    engine = DataParallelEngine(model) if dist.is_available() and dist.is_initialized() else DataParallelEngine(model, world_size=1)
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    graphed = GraphedTrainStep(model, criterion, optimizer, engine) if cuda_graph else None
    history = {"epoch_avg_loss": [], "epoch_avg_gen_loss": [], "val_mae": []}
    ckpt_dir = os.path.join(save_path, "checkpoints") if save_path else ""
    if ckpt_dir and rank == 0:
        os.makedirs(ckpt_dir, exist_ok=True)

    for epoch in range(start_epoch, epochs):
        model.train(True)
        model.set_training(True)
        loss_sum = torch.zeros((), device=device)
        gen_sum = torch.zeros((), device=device)
        num_samples = 0
        for batch in train_loader:
            mri, tau, roi, _, covars, paths = _unpack(batch)
            mri, tau, roi = mri.to(device, non_blocking=True), tau.to(device, non_blocking=True), roi.to(device, non_blocking=True)
            if graphed is not None:      # replay; a ragged last batch runs the same step eagerly on the runner's stream
                step_fn = graphed if graphed.matches(mri) else graphed.eager
                loss = step_fn(mri, tau, roi, covars, roi_pred_fn(paths))
                loss_sum += loss
                gen_sum += graphed.gen_loss.sum()
                num_samples += mri.shape[0]
                continue
            optimizer.zero_grad(set_to_none=True)
            pred, projected, final_repr = model(mri, covars, roi_pred_dicts=roi_pred_fn(paths), sample_roi_mask=roi)[:3]
            feats, labels = engine.gather_rnc(projected[-1], covars[:, -1].float().to(device))       # :842-845
            zeros = torch.zeros(final_repr.size(), device=device)
            loss, gen_loss, _, _ = criterion(pred, tau, roi, (final_repr, zeros, zeros), (feats, labels))   # :878
            loss.backward()
            engine.finish()
            optimizer.step()
            loss_sum += loss.detach()
            gen_sum += gen_loss.detach().sum()
            num_samples += mri.shape[0]
        # the only host sync of the epoch; summed over ranks so every replica's scheduler sees the global-batch figures.  (With
        # several ranks every rank's loss already contains the global RnC term; the generative part is rank-local.)
        epoch_loss, epoch_gen, num_samples = engine.all_reduce_scalars(float(loss_sum), float(gen_sum), float(num_samples))
        scheduler.step(epoch_loss / max(num_samples, 1))
        history["epoch_avg_loss"].append(epoch_loss / max(num_samples, 1))
        history["epoch_avg_gen_loss"].append(epoch_gen / max(num_samples, 1))
        if ckpt_dir and rank == 0:
            ckpt = {"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                    "loss": loss.detach() if num_samples else None, "scheduler_state_dict": scheduler.state_dict()}
            torch.save(ckpt, os.path.join(ckpt_dir, "checkpoint_latest_epoch.pth"))
            if epoch % checkpoint_iter == 0:
                torch.save(ckpt, os.path.join(ckpt_dir, f"checkpoint_epoch_{epoch}.pth"))
        if validation_loader is not None and (epoch % val_iter == 0 or epoch == epochs - 1):
            history["val_mae"].append((epoch, validate(model, validation_loader, roi_pred_fn, device)))
    return history
