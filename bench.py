#!/usr/bin/env python
"""Benchmark of the hot path: 3-D volumes/sec of the CoMA attention U-Net on B200.

    python bench.py --gpus N --steps K --warmup W [--mode both|train|infer] [--config c4|c5 --batch B] [--impl reference]

Default (what the driver runs): the headline `value` is the TRAINING step of BASELINE.json configs[2] -- forward + loss + backward +
bucketed NCCL gradient all-reduce + AdamW, batch 4 per GPU, 1x128^3, bf16 -- and the same JSON line carries an `infer` sub-record for
configs[1] (inference, batch 8 x 1x128^3 per GPU, no data-path collective) with its own `value`, `e2e` and `roofline`.  N>1
(torchrun, one rank per GPU): every rank runs its own batch (weak scaling).

One JSON line on stdout (rank 0).  `value`: device-timed (CUDA events, max over ranks), inputs resident in HBM.  The timed window is
at least `--min-seconds` (3 s) long: when K steps would be shorter the loop runs more steps and `steps` reports the number actually
timed (`steps_requested` = K), so the number is a sustained one, not a boost-clock burst.  `e2e`: the same metric through the public
streaming API with pinned HOST buffers, H2D of every step's inputs and D2H of its result inside the timed region.  `roofline`:
tcgen05 conv kernel family, ALGORITHMIC FLOPs (the layers' true channel counts, zero padding not counted) / CUDA-event time of its
launches in one instrumented step.  `comm`: gradient bytes all-reduced per step and the exposed (non-overlapped) all-reduce time.
`cpu_baseline` / `--impl reference`: the fp32 oracle port of the reference (forward + backward + AdamW for the training metric) on
this box's host cores.

Other workloads (one line each, committed under profiles/): `--config c4 --batch B` = full-resolution 160x192x160 inference with a
GLOBAL batch B split over the N ranks (BASELINE configs[3]); `--config c5` = training on 256^3 crops with a mixed float32/float64
covariate batch (configs[4]).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHANNELS = [32, 64, 128, 256, 512]
FWD_GFLOP_128 = 826.6            # convs only, single backbone pass, per 128^3 volume
TRAIN_GFLOP_128 = 2479.7

WORKLOADS = {
    # name: (kind, shape, default per-GPU batch, description)
    "c2": ("infer", (128, 128, 128), 8, "inference, batch %d x 1x128^3 MRI->tau-PET per GPU (BASELINE configs[1])"),
    "c3": ("train", (128, 128, 128), 4, "training step fwd+loss+bwd+all-reduce+AdamW, batch %d per GPU, 1x128^3 (BASELINE configs[2])"),
    "c4": ("infer", (160, 192, 160), 8, "full-resolution inference 1x160x192x160 (VolumeDataset_Inference shape), GLOBAL batch %d "
                                        "split over the ranks (BASELINE configs[3])"),
    "c5": ("train", (256, 256, 256), 2, "training step on 1x256^3 crops, batch %d per GPU mixing ADNI-shaped (float32 covariates) and "
                                        "A4-shaped (float64 covariates) samples (BASELINE configs[4])"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops_sustained"], "tflops_burst": p["bf16_tflops"],
                "sm_mhz_sustained": (p.get("clocks_under_load") or {}).get("sm_mhz_median"),
                "source": "measured (MEASURED_PEAKS.json: sustained cuBLAS bf16, burst in frac_of_burst)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1400.0, "sm_mhz_sustained": None,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [l for t, l in self.lines if self.t0 is not None and self.t1 is not None and self.t0 <= t <= self.t1 + 0.05]
        window = "timed region"
        if not inside:
            inside = [l for t, l in self.lines if self.t0 is None or t <= (self.t1 or t) + 0.05][-4:]
            window = "warm-up + timed region (timed region shorter than the sampling period)"
        for line in inside:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                if len(parts) > 6:
                    pw.append(float(parts[6]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None, "window": window}


def make_batch(batch, seed, shape=(128, 128, 128), device=None, pin=False, mixed=False):
    """(mri, tau, roi, covars, roi dicts) with the VolumeDataset tuple layout; ``mixed``: odd samples are ADNI-shaped (float32
    covariates, VolumeDataset.py:427), even ones A4/combined-shaped (float64, VolumeDataset_ADNI_A4_combined.py:86) -- stacking
    promotes the batch to float64, which is what the default collate hands the model."""
    from coma_unet_b200 import SyntheticVolumeDataset
    ds64 = SyntheticVolumeDataset(length=batch, shape=shape, seed=seed)
    ds32 = SyntheticVolumeDataset(length=batch, shape=shape, seed=seed, flavour="adni") if mixed else ds64
    items = [(ds32 if (mixed and i % 2) else ds64)[i] for i in range(batch)]
    mri = torch.stack([it[0] for it in items])
    tau = torch.stack([it[1] for it in items])
    roi = torch.stack([it[2] for it in items])
    covars = torch.stack([it[3][1].to(torch.float64) for it in items]) if mixed else torch.stack([it[3][1] for it in items])
    dicts = [ds64.roi_predictions(i) for i in range(batch)]
    if pin:
        mri, tau, roi, covars = mri.pin_memory(), tau.pin_memory(), roi.pin_memory(), covars.pin_memory()
    if device is not None:
        mri, tau, roi = mri.to(device), tau.to(device), roi.to(device)
    return mri, tau, roi, covars, dicts


COMPUTE = {"bf16": (torch.bfloat16, False, "bf16"), "fp32": (torch.float32, False, "f32"), "fp32_tc": (torch.float32, True, "f32 storage, bf16 hi/lo split operands on tcgen05")}


def build_model(device, shape=(128, 128, 128), dtype=torch.bfloat16, seed=0, fp32_tc=False):
    import coma_unet_b200 as cu
    torch.manual_seed(seed)
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, CHANNELS, [2] * 5, latent_spaces=[2048] * 5, conditional=True,
                                       decoder_ds=False, compute_dtype=dtype, prompt_shape=tuple(shape), fp32_tensor_cores=fp32_tc)
    with torch.no_grad():   # random-init weights; make the zero-initialised FiLM layers non-trivial
        for name, p in m.named_parameters():
            if ".film.2." in name:
                p.normal_(0, 0.02)
    m.set_save_attn(None)
    return m.to(device)


def build_criterion(mod=None):
    if mod is None:
        import coma_unet_b200 as mod
        from coma_unet_b200.model import ROI_INDICES
    else:
        from tests.golden.common import ROI_INDICES
    gen = mod.RoiMSE(torch.tensor([225.0] * 36), ROI_INDICES, voxel_wise=False)
    crit = mod.GenerativeContrastiveLoss(mod.RnCLoss(), gen, torch.nn.TripletMarginLoss(1), 0., 1.)
    crit.gen_loss.batch_reduction = None
    return crit


def dist_setup(n):
    if n <= 1:
        return 0, 1, 0
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(v, world, device):
    if world <= 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(v, world, device):
    if world <= 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
# CPU arm: the fp32 oracle port of the reference on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_volumes_per_sec(kind, budget_s=25.0, steps=1, warmup=0):
    """fp32 oracle (restatement of the reference; the reference itself is not importable: no monai / CondConv) on this box's
    host cores, one 1x128^3 volume per step.  kind "train": forward (incl. the reference's duplicated backbone pass) + loss +
    backward + AdamW.  Returns (volumes/s, cores, sample description)."""
    from oracle import criterions as ocrit
    from oracle import model as omodel
    from tests.golden import common
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shape = (128, 128, 128)
    model = omodel.ContrastiveAttentionUNET_DP(3, 1, 1, CHANNELS, [2] * 5, latent_spaces=[2048] * 5, conditional=True,
                                               prompt_shape=shape)
    mri, tau, roi, covars, dicts = common.synthetic_batch(1, shape, 1234)
    train = kind == "train"
    if train:
        model.train(True)
        model.set_training(True)
        crit = build_criterion(ocrit)
        opt = torch.optim.AdamW(model.parameters(), 1e-3)
    else:
        model.eval()
        model.set_training(False)
    times = []
    t_all = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if train:
            opt.zero_grad(set_to_none=True)
            pred, proj, final = model(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
            z = torch.zeros(final.size())
            loss, _, _, _ = crit(pred, tau, roi, (final, z, z), (proj[-1], covars[:, -1].float()))
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                model(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_all > budget_s and times:
            break
    per = sum(times) / len(times)
    what = "forward + loss + backward + AdamW" if train else "inference"
    return 1.0 / per, cores, (f"{len(times)} x (1 volume 1x128^3, {what}, fp32 oracle port incl. the reference's duplicated "
                              f"backbone pass, {cores} threads)")


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    kind = "infer" if args.mode == "infer" else "train"
    t_start = time.perf_counter()
    vps, cores, sample = cpu_oracle_volumes_per_sec(kind, budget_s=150.0, steps=max(args.steps, 1), warmup=min(max(args.warmup, 0), 1))
    wl = WORKLOADS["c3" if kind == "train" else "c2"]
    line = {
        "impl": "reference", "metric": "volumes_per_sec_" + kind, "value": vps, "unit": "volumes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / vps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (wl[3] % wl[2]) + "; the CPU arm times a bounded sample: 1 volume per step"},
        "cpu_baseline": {"value": vps, "unit": "volumes/s", "cores": cores, "kind": "port",
                         "sample": sample + "; the reference is not importable (monai / CondConv missing), wall %.0f s"
                                   % (time.perf_counter() - t_start)},
        "e2e": {"value": vps, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# instrumented pass: CUDA events around every ABI call
# ------------------------------------------------------------------------------------------------
CONV_CALLS = ("coma_conv3d_fprop", "coma_convT3d_fprop", "coma_conv3d_dgrad", "coma_convT3d_dgrad")
WGRAD_CALLS = ("coma_conv3d_wgrad", "coma_convT3d_wgrad")
HBM_CALLS = ("coma_gate_fwd", "coma_norm_film_act_fwd", "coma_norm_film_act_bwd", "coma_gate_apply_fwd", "coma_roi_paint",
             "coma_pack2_fwd", "coma_norm_stats")


def kernel_profile(step_fn, n=2):
    """-> (families, conv layer list, hbm kernels, timeline).  FLOPs are ALGORITHMIC: the layer's true channel counts
    (`alg`, set by coma_unet_b200.ops), not the zero-padded ones the kernel computes on; `flops_executed` keeps the padded figure."""
    from coma_unet_b200 import _lib
    records = []
    orig = _lib.call

    def timed(name, *cargs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *cargs)
        e1.record()
        flops = executed = nbytes = 0.0
        tc, shape = False, ()
        if name in CONV_CALLS:
            a = cargs[0]._obj
            vox = a.Di * a.Hi * a.Wi if a.transposed else a.Do * a.Ho * a.Wo
            cin, cout = getattr(a, "alg", (a.Cin, a.Cout))
            flops = 2.0 * a.B * vox * a.ksize ** 3 * cin * cout
            executed = 2.0 * a.B * vox * a.ksize ** 3 * a.Cin * a.Cout
            tc = _lib.lib().coma_conv3d_impl(cargs[0]) == _lib.IMPL_TCGEN05
            shape = (a.B, cin, cout, a.Do, a.Ho, a.Wo, a.ksize, a.stride, a.transposed)
        elif name in WGRAD_CALLS:
            a = cargs[0]._obj
            cg, cx = getattr(a, "alg", (a.Cg, a.Cx))
            flops = 2.0 * a.B * a.Dg * a.Hg * a.Wg * a.ksize ** 3 * cg * cx
            executed = 2.0 * a.B * a.Dg * a.Hg * a.Wg * a.ksize ** 3 * a.Cg * a.Cx
            tc = bool(_lib.lib().coma_conv3d_wgrad_tcgen05_supported(cargs[0]))
            shape = (a.B, cg, cx, a.Dg, a.Hg, a.Wg, a.ksize, a.stride, 0)
        elif name in HBM_CALLS and name != "coma_norm_stats":
            a = cargs[0]._obj
            es = 2 if a.dtype == _lib.BF16 else 4
            vox = float(a.B) * float(a.V)
            if name == "coma_gate_fwd":
                nbytes = 3.0 * a.C * es * vox                      # read g, read x, write out
            elif name == "coma_norm_film_act_fwd":
                nbytes = (2.0 + (1.0 if a.r else 0.0)) * a.C * es * vox
            elif name == "coma_norm_film_act_bwd":                 # reduce sweep reads x, dy; apply sweep reads x, dy, writes dx
                nbytes = (5.0 + (1.0 if a.r else 0.0) + (1.0 if a.dr else 0.0)) * a.C * es * vox
            elif name == "coma_gate_apply_fwd":
                nbytes = (2.0 * a.C + 1.0) * es * vox
            elif name == "coma_roi_paint":
                nbytes = (8.0 + 4.0 + a.out_cs * es) * vox         # roi + mri (fp32) + prompt, write out_cs channels
            else:
                nbytes = (2.0 * es + a.dst_cs * es) * vox
        records.append((name, tc, flops, executed, nbytes, e0, e1, shape))

    _lib.call = timed
    try:
        for _ in range(n):
            records.clear()
            step_fn()
            torch.cuda.synchronize()
    finally:
        _lib.call = orig
    fam, layers, hbm, timeline = {}, [], {}, []
    for name, tc, flops, executed, nbytes, e0, e1, shape in records:
        ms = e0.elapsed_time(e1)
        key = name + (":tcgen05" if tc else "")
        f = fam.setdefault(key, {"ms": 0.0, "flops": 0.0, "flops_executed": 0.0, "launches": 0})
        f["ms"] += ms
        f["flops"] += flops
        f["flops_executed"] += executed
        f["launches"] += 1
        if flops:
            layers.append({"kernel": key, "B,Cin,Cout,Do,Ho,Wo,k,stride,T": list(shape), "ms": ms, "flops": flops,
                           "flops_executed": executed, "tflops": flops / ms / 1e9})
        if nbytes:
            h = hbm.setdefault(name, {"ms": 0.0, "bytes": 0.0, "launches": 0, "best_large_launch_GBps": 0.0})
            h["ms"] += ms
            h["bytes"] += nbytes
            h["launches"] += 1
            if nbytes > 2.5e8:                        # launches big enough to be bandwidth- rather than latency-bound
                h["best_large_launch_GBps"] = max(h["best_large_launch_GBps"], nbytes / ms / 1e6)
        timeline.append({"call": key, "ms": ms, "shape": list(shape) if shape else None})
    return fam, layers, hbm, timeline


def roofline_records(fam, layers, hbm, peaks, clocks, traffic_file):
    tc = [v for k, v in fam.items() if k.endswith(":tcgen05")]
    tc_ms = sum(v["ms"] for v in tc)
    tc_flops = sum(v["flops"] for v in tc)
    tc_exec = sum(v["flops_executed"] for v in tc)
    tc_n = sum(v["launches"] for v in tc)
    total_ms = sum(v["ms"] for v in fam.values())
    achieved = tc_flops / (tc_ms / 1000.0) / 1e12 if tc_ms > 0 else 0.0
    top = max((l for l in layers if l["kernel"].endswith(":tcgen05")), key=lambda l: l["ms"], default=None)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", traffic_file)) as f:
            for k in json.load(f)["kernels"]:
                if top is not None and k.get("match") == top["B,Cin,Cout,Do,Ho,Wo,k,stride,T"][:4] + [top["kernel"].split(":")[0]]:
                    traffic, traffic_src = k["dram_traffic_bytes"], f"profiles/{traffic_file} (ncu --set full, dram read+write, one launch)"
    except (OSError, KeyError, IndexError, ValueError, TypeError):
        pass
    roofline = {
        "bound": "tensor",
        "kernel": "tcgen05 implicit-GEMM conv family (conv_halo3 / conv_pair / conv_halo_s2 / convT_halo / conv_tc, in training also "
                  "their dgrad launches and wgrad_tc; all launches of one step)",
        "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
        "frac_of_burst_peak": achieved / peaks["tflops_burst"], "peak_burst": peaks["tflops_burst"],
        "flops_counted": "algorithmic: 2*taps*Cin*Cout per voxel with the layer's true channel counts (zero padding to 16 channels is "
                         "not counted)",
        "achieved_incl_padding": tc_exec / (tc_ms / 1000.0) / 1e12 if tc_ms > 0 else 0.0,
        "sm_mhz_window": clocks.get("sm_mhz"), "sm_mhz_of_sustained_peak": peaks.get("sm_mhz_sustained"),
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"], "launches_per_step": tc_n,
        "avg_launch_ms": tc_ms / max(tc_n, 1), "share_of_step": tc_ms / max(total_ms, 1e-9), "family_gflop_per_step": tc_flops / 1e9,
        "top_launch": None if top is None else {
            "kernel": top["kernel"], "B,Cin,Cout,Do,Ho,Wo,k,stride,T": top["B,Cin,Cout,Do,Ho,Wo,k,stride,T"], "ms": top["ms"],
            "achieved": top["tflops"], "frac": top["tflops"] / peaks["tflops"], "frac_of_burst_peak": top["tflops"] / peaks["tflops_burst"],
            "algorithmic_flops": top["flops"]}}
    hbm_roof = {k: {"ms": v["ms"], "algorithmic_GB": v["bytes"] / 1e9, "launches": v["launches"],
                    "achieved_GBps": v["bytes"] / v["ms"] / 1e6, "frac_of_measured_peak": v["bytes"] / v["ms"] / 1e6 / peaks["hbm_gbs"],
                    "best_large_launch_GBps": v["best_large_launch_GBps"],
                    "best_large_launch_frac": v["best_large_launch_GBps"] / peaks["hbm_gbs"]}
                for k, v in hbm.items() if v["ms"] > 0}
    return roofline, hbm_roof


# ------------------------------------------------------------------------------------------------
def run_workload(name, args, rank, world, local, device, peaks, batch_override=0):
    """Time one workload; returns the record (all ranks compute it, rank 0 prints)."""
    from coma_unet_b200 import DevicePrefetcher, HostSink, _lib
    kind, shape, default_batch, text = WORKLOADS[name]
    train = kind == "train"
    voxel_scale = shape[0] * shape[1] * shape[2] / 128.0 ** 3
    if name == "c4":                     # a GLOBAL batch split over the ranks (ranks beyond the batch stay idle)
        global_batch = batch_override or default_batch
        per = (global_batch + world - 1) // world
        batch = max(0, min(per, global_batch - rank * per))
    else:
        batch = batch_override or default_batch
        global_batch = batch * world
    torch.cuda.reset_peak_memory_stats(device)
    cdtype, fp32_tc, dtype_label = COMPUTE[args.compute]
    model = build_model(device, shape, dtype=cdtype, fp32_tc=fp32_tc)
    nb = max(batch, 1)
    mri, tau, roi, covars, dicts = make_batch(nb, 1234 + rank, shape, device=device, mixed=(name == "c5"))
    h_mri, h_tau, h_roi, h_cov, _ = make_batch(nb, 1234 + rank, shape, pin=True, mixed=(name == "c5"))
    engine = None
    use_graph = not args.no_graph
    if train:
        from coma_unet_b200.graph import GraphedTrainStep
        from coma_unet_b200.parallel import DataParallelEngine
        model.train(True)
        crit = build_criterion()
        engine = DataParallelEngine(model, world_size=world)
        # the optimizer train_dp builds: fused AdamW (capturable + tensor learning rate when the step is replayed from a CUDA graph)
        opt = GraphedTrainStep.make_optimizer(model, 1e-3) if use_graph else torch.optim.AdamW(model.parameters(), 1e-3, fused=True)
        graphed = GraphedTrainStep(model, crit, opt, engine) if use_graph else None

        def eager_step(m=mri, t=tau, r=roi, c=covars):
            opt.zero_grad(set_to_none=True)
            pred, proj, final = model(m, c, roi_pred_dicts=dicts, sample_roi_mask=r)
            feats, labels = engine.gather_rnc(proj[-1], c[:, -1].float().to(device, non_blocking=True))
            z = torch.zeros(final.size(), device=device)
            loss, gen, _, _ = crit(pred, t, r, (final, z, z), (feats, labels))
            loss.backward()
            engine.finish()
            opt.step()
            return loss

        def step(m=mri, t=tau, r=roi, c=covars):
            return graphed(m, t, r, c, dicts) if use_graph else eager_step(m, t, r, c)
    else:
        from coma_unet_b200.graph import GraphedInference
        model.eval()
        model.set_training(False)
        graphed = GraphedInference(model) if use_graph and batch else None

        def eager_step(m=mri, t=tau, r=roi, c=covars):
            if batch == 0:
                return None
            with torch.no_grad():
                return model(m, c, roi_pred_dicts=dicts, sample_roi_mask=r)

        def step(m=mri, t=tau, r=roi, c=covars):
            return graphed(m, c, dicts, r) if graphed is not None else eager_step(m, t, r, c)

    def timed_steps(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        return e0, e1

    warm = max(args.warmup, 5 if use_graph else 3)       # graph replay: 2 eager steps + the capture come first
    with ClockSampler(local) as clocks:           # sampling starts with the warm-up (same load) and is windowed to the timed region
        for _ in range(warm - 2):
            step()
        torch.cuda.synchronize()
        e0, e1 = timed_steps(2)                   # the last two warm-up steps calibrate the length of the timed window
        torch.cuda.synchronize()
        est_ms = max_over_ranks(e0.elapsed_time(e1) / 2.0, world, device)
        steps = max(args.steps, int(math.ceil(1.08 * args.min_seconds * 1000.0 / max(est_ms, 1e-3))))
        barrier(world)
        torch.cuda.synchronize()
        l0 = _lib.launches
        clocks.mark_start()
        e0, e1 = timed_steps(steps)
        barrier(world)
        torch.cuda.synchronize()
        clocks.mark_end()
        ms = max_over_ranks(e0.elapsed_time(e1), world, device)
    launches = _lib.launches - l0
    value = global_batch * steps / (ms / 1000.0)
    clk = clocks.summary()

    # ---- communication (training, N > 1): bytes per step and the exposed all-reduce tail ----
    comm = None
    if train:
        comm = {"allreduce_bytes_per_step": 0, "collectives": "none (single rank)"}
        if engine.enabled:
            engine.timing, engine.timing_events = True, []
            for _ in range(10):
                eager_step()
            torch.cuda.synchronize()
            engine.timing = False
            exposed = [a.elapsed_time(b) for a, b in engine.timing_events]
            comm = {"allreduce_bytes_per_step": engine.comm_bytes, "buckets": len(engine.launch_log),
                    "buckets_launched_during_backward": engine.launched_in_backward,
                    "late_buckets": len(engine.buckets) - engine.n_early,
                    "exposed_allreduce_ms": max_over_ranks(sum(exposed) / max(len(exposed), 1), world, device),
                    "exposed_allreduce_ms_this_rank_max": max(exposed) if exposed else 0.0,
                    "how": "CUDA events on the compute stream around its wait for the NCCL stream at the end of backward (mean of 10 "
                           "steps, max over ranks); gradient SUM all-reduce in ~25 MB flat buckets launched from autograd hooks in a fixed "
                           "order + all-gather of the [B,512] RnC features and [B,6] labels + a host-side (gloo) MAX of the used-parameter mask",
                    "rnc_allgather_bytes_per_step": int(world * batch * (512 + 6) * 4)}

    # ---- end to end: pinned host inputs -> H2D -> model -> D2H of the result, every step ----
    h_out = torch.empty((nb, 1, *shape), dtype=torch.float32).pin_memory() if not train else torch.empty(1).pin_memory()
    h2d = (h_mri.numel() * 4 + h_roi.numel() * 4 + (h_tau.numel() * 4 if train else 0) + h_cov.numel() * h_cov.element_size()) if batch else 0
    d2h = h_out.numel() * 4 if batch else 0
    # The public streaming API (coma_unet_b200.DevicePrefetcher / HostSink): every step's inputs are copied from pinned host memory
    # and every step's result is read back to pinned host memory inside the timed region; the copies run on their own streams, one
    # batch ahead / behind the compute stream.
    sink = HostSink(h_out.shape, torch.float32, device)

    def host_batches(n):
        for _ in range(n):
            yield (h_mri, h_tau, h_roi) if train else (h_mri, h_roi)

    def e2e_run(n):
        if batch == 0:
            return
        for dev in DevicePrefetcher(host_batches(n), device):
            if train:
                m, t, r = dev
                sink.put(step(m, t, r, h_cov))
            else:
                m, r = dev
                sink.put(step(m, None, r, h_cov))
        sink.wait()

    e2e_run(2)
    torch.cuda.synchronize()
    barrier(world)
    t0 = time.perf_counter()
    e0.record()
    e2e_run(steps)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    barrier(world)
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms), world, device)
    e2e_value = global_batch * steps / (e2e_ms / 1000.0)

    # ---- roofline of the dominant kernel family (instrumented extra pass, not part of the timing above) ----
    roofline = hbm_roof = None
    if batch:
        fam, layers, hbm, timeline = kernel_profile(eager_step)
        roofline, hbm_roof = roofline_records(fam, layers, hbm, peaks, clk, "r02_ncu_conv_kernels_summary.json")
        if rank == 0 and args.profile_out:
            with open(args.profile_out.replace(".json", f"_{name}.json") if args.mode == "both" and not args.config else args.profile_out, "w") as f:
                json.dump({"workload": text % (global_batch if name == "c4" else batch), "hbm_bound_kernels": hbm_roof,
                           "families": fam, "conv_layers": layers, "timeline": timeline}, f, indent=1)
    peak_mem = max_over_ranks(torch.cuda.max_memory_allocated(device) / 2 ** 30, world, device)

    gflop = (TRAIN_GFLOP_128 if train else FWD_GFLOP_128) * voxel_scale
    per_rank = max((global_batch + world - 1) // world, 1)
    n_active = world if name != "c4" else min(world, (global_batch + per_rank - 1) // per_rank)
    rec = {
        "metric": "volumes_per_sec_" + kind, "value": value, "unit": "volumes/s", "n_gpus": world, "steps": steps,
        "steps_requested": args.steps, "warmup": warm, "ms_per_step": ms / steps, "timed_window_s": ms / 1000.0,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_label, "data": "synthetic",
        "config": {"workload": text % (global_batch if name == "c4" else batch), "config": name, "shape": list(shape),
                   "channels": CHANNELS, "per_gpu_batch": batch if name != "c4" else (global_batch + world - 1) // world,
                   "global_batch": global_batch, "parallelism": f"dp{world}", "active_gpus": n_active,
                   "l2": "activations per step (>10 GB) exceed the 126 MB L2; no explicit flush",
                   "launch": "CUDA graph replay (coma_unet_b200.graph), one graph per step" if use_graph else "eager (one ABI call per kernel)",
                   "gflop_per_volume_algorithmic": gflop, "model_tflops": value * gflop / 1e3,
                   "model_frac_of_peak": value * gflop / 1e3 / (peaks["tflops"] * max(n_active, 1)),
                   "peak_hbm_allocated_GiB": peak_mem},
        "roofline": roofline, "roofline_hbm_kernels": hbm_roof, "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
    }
    if comm is not None:
        rec["comm"] = comm
    del model, mri, tau, roi, h_mri, h_tau, h_roi, sink, graphed, step, eager_step
    if train:
        del opt, engine
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    rank, world, local = dist_setup(args.gpus)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    peaks = load_peaks()
    if args.config:
        names = [args.config]
    else:
        names = {"both": ["c3", "c2"], "train": ["c3"], "infer": ["c2"]}[args.mode]
    recs = [run_workload(n, args, rank, world, local, device, peaks, batch_override=args.batch) for n in names]
    line = recs[0]
    if len(recs) > 1:
        sub = recs[1]
        for k in ("n_gpus", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "unit"):
            sub.pop(k, None)
        line["infer"] = sub
    line["cpu_baseline"] = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = WORKLOADS[names[0]][0]
        vps, cores, sample = cpu_oracle_volumes_per_sec(kind, budget_s=25.0, steps=3, warmup=0)
        line["cpu_baseline"] = {"value": vps, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="both", choices=["both", "train", "infer"],
                    help="both (default): training headline (configs[2]) + `infer` sub-record (configs[1])")
    ap.add_argument("--config", default="", choices=["", "c2", "c3", "c4", "c5"], help="time one named BASELINE config instead")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (c4: the GLOBAL batch split over the ranks)")
    ap.add_argument("--min-seconds", type=float, default=3.0, help="minimum length of the timed window")
    ap.add_argument("--compute", default="bf16", choices=list(COMPUTE), help="bf16 (the benchmarked path), fp32 (exact, CUDA-core convs), "
                    "fp32_tc (fp32 storage, 3x3x3 convs on tcgen05 through split-precision operands)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--profile-out", default="")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
