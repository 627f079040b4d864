#!/usr/bin/env python
"""Benchmark of the hot path: 3-D volumes/sec of the CoMA attention U-Net on B200.

    python bench.py --gpus N --steps K --warmup W [--mode infer|train] [--impl reference]

Default workload (N=1) = BASELINE.json configs[1]: inference, batch 8, synthetic 1x128^3 MRI -> tau-PET,
bf16, one B200.  N>1 (torchrun, one rank per GPU): every rank runs its own batch (weak scaling, no data-path
collective for inference; --mode train adds the bucketed NCCL gradient all-reduce).
One JSON line on stdout (rank 0).  `value`: device-timed, inputs resident in HBM.  `e2e`: the same metric through
the public model call with pinned HOST buffers, H2D of the inputs and D2H of the prediction inside the timed
region.  `roofline`: tcgen05 conv kernel family, algorithmic FLOPs / CUDA-event time of its launches in one step.
`cpu_baseline` / `--impl reference`: the fp32 oracle port of the reference on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHANNELS = [32, 64, 128, 256, 512]
SHAPE = (128, 128, 128)
FWD_GFLOP_PER_VOLUME = 826.6          # convs only, single backbone pass, 128^3
TRAIN_GFLOP_PER_VOLUME = 2479.7


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops_sustained"], "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # samples taken while the timed region ran; a region shorter than the sampling period borrows the samples of the
        # warm-up steps right before it (the sampler starts with the warm-up, same load)
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
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "window": window}


def make_batch(batch, seed, device=None, pin=False):
    from coma_unet_b200 import SyntheticVolumeDataset
    ds = SyntheticVolumeDataset(length=batch, shape=SHAPE, seed=seed)
    items = [ds[i] for i in range(batch)]
    mri = torch.stack([it[0] for it in items])
    tau = torch.stack([it[1] for it in items])
    roi = torch.stack([it[2] for it in items])
    covars = torch.stack([it[3][1] for it in items])
    dicts = [ds.roi_predictions(i) for i in range(batch)]
    if pin:
        mri, tau, roi, covars = mri.pin_memory(), tau.pin_memory(), roi.pin_memory(), covars.pin_memory()
    if device is not None:
        mri, tau, roi = mri.to(device), tau.to(device), roi.to(device)
    return mri, tau, roi, covars, dicts


def build_model(device, dtype=torch.bfloat16, seed=0):
    import coma_unet_b200 as cu
    torch.manual_seed(seed)
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, CHANNELS, [2] * 5, latent_spaces=[2048] * 5, conditional=True,
                                       decoder_ds=False, compute_dtype=dtype)
    with torch.no_grad():   # random-init weights; make the zero-initialised FiLM layers non-trivial
        for name, p in m.named_parameters():
            if ".film.2." in name:
                p.normal_(0, 0.02)
    m.set_save_attn(None)
    return m.to(device)


def build_criterion():
    import coma_unet_b200 as cu
    from coma_unet_b200.model import ROI_INDICES
    gen = cu.RoiMSE(torch.tensor([225.0] * 36), ROI_INDICES, voxel_wise=False)
    crit = cu.GenerativeContrastiveLoss(cu.RnCLoss(), gen, torch.nn.TripletMarginLoss(1), 0., 1.)
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


def max_over_ranks(ms, world, device):
    if world <= 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
def cpu_oracle_volumes_per_sec(budget_s=25.0, steps=1, warmup=0):
    """fp32 oracle (restatement of the reference; the reference itself is not importable: no monai / CondConv)
    on this box's host cores.  Returns (volumes/s, cores, sample description)."""
    from oracle import model as omodel
    from tests.golden import common
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shape = SHAPE
    model = omodel.ContrastiveAttentionUNET_DP(3, 1, 1, CHANNELS, [2] * 5, latent_spaces=[2048] * 5, conditional=True,
                                               prompt_shape=shape).eval()
    model.set_training(False)
    mri, tau, roi, covars, dicts = common.synthetic_batch(1, shape, 1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            model(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if sum(times) > budget_s and times:
                break
    per = sum(times) / len(times)
    return 1.0 / per, cores, f"{len(times)} x (1 volume 1x128^3, inference, fp32 oracle incl. the reference's duplicated backbone pass)"


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    t_start = time.perf_counter()
    budget = 200.0
    vps, cores, _ = cpu_oracle_volumes_per_sec(budget_s=budget, steps=max(args.steps, 1) + max(args.warmup, 0), warmup=0)
    # the first `warmup` iterations are part of the measured sample only if the budget cut the run short
    line = {
        "impl": "reference", "metric": "volumes_per_sec_infer", "value": vps, "unit": "volumes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / vps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "inference, batch 8 x 1x128^3 MRI->tau-PET (BASELINE configs[1]); CPU arm times 1 volume per step"},
        "cpu_baseline": {"value": vps, "unit": "volumes/s", "cores": cores, "kind": "port",
                         "sample": "1 volume (1x128^3) per step, fp32 oracle port of the reference (reference not importable: "
                                   "monai/CondConv missing), all host threads, wall %.0f s" % (time.perf_counter() - t_start)},
        "e2e": {"value": vps, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def conv_flops_of_call(name, a):
    taps = a.ksize ** 3
    if a.transposed:
        vox = a.Di * a.Hi * a.Wi
    else:
        vox = a.Do * a.Ho * a.Wo
    return 2.0 * a.B * vox * taps * a.Cin * a.Cout


def kernel_profile(step_fn, n=2):
    """One instrumented pass: CUDA events around every ABI call -> per-family time and conv FLOPs."""
    from coma_unet_b200 import _lib
    records = []
    orig = _lib.call

    def timed(name, *cargs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *cargs)
        e1.record()
        flops, tc = 0.0, False
        if name in ("coma_conv3d_fprop", "coma_convT3d_fprop", "coma_conv3d_dgrad", "coma_convT3d_dgrad"):
            a = cargs[0]._obj
            flops = conv_flops_of_call(name, a)
            tc = _lib.lib().coma_conv3d_impl(cargs[0]) == _lib.IMPL_TCGEN05
            shape = (a.B, a.Cin, a.Cout, a.Do, a.ksize, a.stride, a.transposed)
        elif name in ("coma_gate_fwd", "coma_norm_film_act_fwd", "coma_gate_apply_fwd", "coma_roi_paint", "coma_pack2_fwd"):
            a = cargs[0]._obj
            es = 2 if a.dtype == _lib.BF16 else 4
            vox = float(a.B) * float(a.V)
            if name == "coma_gate_fwd":
                nbytes = 3.0 * a.C * es * vox                      # read g, read x, write out
            elif name == "coma_norm_film_act_fwd":
                nbytes = (2.0 + (1.0 if a.r else 0.0)) * a.C * es * vox
            elif name == "coma_gate_apply_fwd":
                nbytes = (2.0 * a.C + 1.0) * es * vox
            elif name == "coma_roi_paint":
                nbytes = (8.0 + 4.0 + a.out_cs * es) * vox         # roi + mri (fp32) + prompt, write out_cs channels
            else:
                nbytes = (2.0 * es + a.dst_cs * es) * vox
            shape = ("bytes", nbytes)
        elif name in ("coma_conv3d_wgrad", "coma_convT3d_wgrad"):
            a = cargs[0]._obj
            flops = 2.0 * a.B * a.Dg * a.Hg * a.Wg * a.ksize ** 3 * a.Cg * a.Cx
            tc = bool(_lib.lib().coma_conv3d_wgrad_tcgen05_supported(cargs[0]))
            shape = (a.B, a.Cg, a.Cx, a.Dg, a.ksize, a.stride, 0)
        else:
            shape = ()
        records.append((name, tc, flops, e0, e1, shape))

    _lib.call = timed
    try:
        for _ in range(n):
            records.clear()
            step_fn()
            torch.cuda.synchronize()
    finally:
        _lib.call = orig
    fam = {}
    layers = []
    hbm = {}
    for name, tc, flops, e0, e1, shape in records:
        ms = e0.elapsed_time(e1)
        key = name + (":tcgen05" if tc else "")
        f = fam.setdefault(key, [0.0, 0.0, 0])
        f[0] += ms
        f[1] += flops
        f[2] += 1
        if flops:
            layers.append((key, shape, ms, flops))
        if shape and shape[0] == "bytes":
            h = hbm.setdefault(name, [0.0, 0.0, 0, 0.0])
            h[0] += ms
            h[1] += shape[1]
            h[2] += 1
            if shape[1] > 2.5e8:                      # launches big enough to be bandwidth- rather than latency-bound
                h[3] = max(h[3], shape[1] / ms / 1e6)
    kernel_profile.hbm = hbm
    kernel_profile.timeline = [{"call": name + (":tcgen05" if tc else ""), "ms": e0.elapsed_time(e1),
                                "shape": list(shape) if shape else None} for name, tc, flops, e0, e1, shape in records]
    return fam, layers


def run_ours(args):
    rank, world, local = dist_setup(args.gpus)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    peaks = load_peaks()
    from coma_unet_b200 import _lib
    train = args.mode == "train"
    batch = args.batch or (4 if train else 8)
    model = build_model(device)
    mri, tau, roi, covars, dicts = make_batch(batch, 1234 + rank, device=device)
    h_mri, h_tau, h_roi, h_cov, _ = make_batch(batch, 1234 + rank, pin=True)

    if train:
        from coma_unet_b200.parallel import DataParallelEngine
        model.train(True)
        crit = build_criterion()
        engine = DataParallelEngine(model, world_size=world)
        opt = torch.optim.AdamW(model.parameters(), 1e-3, fused=True)     # the same optimizer train_dp builds

        def step(m=mri, t=tau, r=roi, c=covars):
            opt.zero_grad(set_to_none=True)
            pred, proj, final = model(m, c, roi_pred_dicts=dicts, sample_roi_mask=r)
            feats, labels = engine.gather_rnc(proj[-1], c[:, -1].float().to(device, non_blocking=True))
            z = torch.zeros(final.size(), device=device)
            loss, gen, _, _ = crit(pred, t, r, (final, z, z), (feats, labels))
            loss.backward()
            engine.finish()
            opt.step()
            return loss
    else:
        model.eval()
        model.set_training(False)

        def step(m=mri, t=tau, r=roi, c=covars):
            with torch.no_grad():
                return model(m, c, roi_pred_dicts=dicts, sample_roi_mask=r)

    with ClockSampler(local) as clocks:           # sampling starts with the warm-up (same load) and is windowed to the timed region
        for _ in range(max(args.warmup, 3)):
            step()
        barrier(world)
        torch.cuda.synchronize()
        l0 = _lib.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks.mark_start()
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        barrier(world)
        torch.cuda.synchronize()
        clocks.mark_end()
        ms = max_over_ranks(e0.elapsed_time(e1), world, device)
    launches = _lib.launches - l0
    value = world * batch * args.steps / (ms / 1000.0)

    # ---- end to end: pinned host inputs -> H2D -> model -> D2H of the result, every step ----
    h_out = torch.empty((batch, 1, *SHAPE), dtype=torch.float32).pin_memory() if not train else torch.empty(1).pin_memory()
    h2d = h_mri.numel() * 4 + h_roi.numel() * 4 + (h_tau.numel() * 4 if train else 0) + h_cov.numel() * 8
    d2h = h_out.numel() * 4

    # The public streaming API (coma_unet_b200.DevicePrefetcher / HostSink): every step's inputs are copied from pinned
    # host memory and every step's result is read back to pinned host memory inside the timed region; the copies run on
    # their own streams, one batch ahead / behind the compute stream.
    from coma_unet_b200 import DevicePrefetcher, HostSink
    sink = HostSink(h_out.shape, torch.float32, device)

    def host_batches(n):
        for _ in range(n):
            yield (h_mri, h_tau, h_roi) if train else (h_mri, h_roi)

    def e2e_run(n):
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
    e2e_run(args.steps)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    barrier(world)
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms), world, device)
    e2e_value = world * batch * args.steps / (e2e_ms / 1000.0)

    # ---- roofline of the dominant kernel family (instrumented extra pass, not part of the timing above) ----
    fam, layers = kernel_profile(step)
    tc_keys = [k for k in fam if k.endswith(":tcgen05")]
    tc_ms = sum(fam[k][0] for k in tc_keys)
    tc_flops = sum(fam[k][1] for k in tc_keys)
    tc_n = sum(fam[k][2] for k in tc_keys)
    total_ms = sum(v[0] for v in fam.values())
    achieved = tc_flops / (tc_ms / 1000.0) / 1e12 if tc_ms > 0 else 0.0
    # the single largest launch of the family, live; its DRAM traffic per launch comes from the committed ncu --set full capture
    top = max((l for l in layers if l[0].endswith(":tcgen05")), key=lambda l: l[2], default=None)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_conv_kernels_summary.json")) as f:
            k0 = json.load(f)["kernels"][0]                  # 64->32 @128^3, the largest launch of the step
        if top is not None and list(top[1])[:4] == [8, 64, 32, 128]:
            traffic, traffic_src = k0["dram_traffic_bytes"], "profiles/r01_ncu_conv_kernels_summary.json (dram read+write, one launch)"
    except (OSError, KeyError, IndexError, ValueError):
        pass
    roofline = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv family (conv_halo3 / conv_pair / conv_halo_s2 / convT_halo / conv_tc kernels and, in training, wgrad_tc; all launches of one step)",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["source"], "launches_per_step": tc_n,
                "avg_launch_ms": tc_ms / max(tc_n, 1), "share_of_step": tc_ms / max(total_ms, 1e-9),
                "top_launch": None if top is None else {
                    "B,Cin,Cout,Do,k,stride,T": list(top[1]), "ms": top[2], "achieved": top[3] / top[2] / 1e9,
                    "frac": top[3] / top[2] / 1e9 / peaks["tflops"], "algorithmic_flops": top[3],
                    "algorithmic_bytes": 2.0 * top[1][0] * top[1][3] ** 3 * (top[1][1] + top[1][2])}}
    hbm_roof = {k: {"ms": v[0], "algorithmic_GB": v[1] / 1e9, "launches": v[2], "achieved_GBps": v[1] / v[0] / 1e6,
                    "frac_of_measured_peak": v[1] / v[0] / 1e6 / peaks["hbm_gbs"],
                    "best_large_launch_GBps": v[3], "best_large_launch_frac": v[3] / peaks["hbm_gbs"]}
                for k, v in getattr(kernel_profile, "hbm", {}).items() if v[0] > 0}
    if rank == 0 and args.profile_out:
        with open(args.profile_out, "w") as f:
            json.dump({"hbm_bound_kernels": hbm_roof,
                       "families": {k: {"ms": v[0], "gflop": v[1] / 1e9, "launches": v[2]} for k, v in fam.items()},
                       "conv_layers": [{"kernel": k, "B,Cin,Cout,Do,k,stride,T": s, "ms": ms_, "tflops": fl / ms_ / 1e9}
                                       for k, s, ms_, fl in layers],
                       "timeline": getattr(kernel_profile, "timeline", [])}, f, indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        vps, cores, sample = cpu_oracle_volumes_per_sec(budget_s=25.0, steps=2, warmup=0)
        cpu = {"value": vps, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        gflop = TRAIN_GFLOP_PER_VOLUME if train else FWD_GFLOP_PER_VOLUME
        line = {
            "metric": "volumes_per_sec_train" if train else "volumes_per_sec_infer", "value": value, "unit": "volumes/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("training step fwd+bwd+AdamW, batch %d per GPU, 1x128^3 (BASELINE configs[2])" % batch) if train
                       else ("inference, batch %d x 1x128^3 MRI->tau-PET per GPU (BASELINE configs[1])" % batch),
                       "channels": CHANNELS, "per_gpu_batch": batch, "parallelism": f"dp{world}",
                       "l2": "activations per step (>10 GB) exceed the 126 MB L2; no explicit flush",
                       "model_tflops": value * gflop / 1e3, "model_frac_of_peak": value * gflop / 1e3 / peaks["tflops"]},
            "roofline": roofline, "roofline_hbm_kernels": hbm_roof, "cpu_baseline": cpu, "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default="")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
