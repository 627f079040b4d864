"""Per-layer error growth of the CUDA path against the fp32 oracle on the same GPU.

    python scripts/error_trace.py [--shape 128] [--width 32] [--seed 7] [--batch 1] [--out gpurun_out/error_trace.json]

Both models get the same deterministic weights (tests/golden/common.py) and inputs; forward hooks on the modules the two trees
share (same names: the state_dict layouts are identical) record every intermediate tensor of an eval forward.  Printed per
module, in execution order: max|a-b| / max|b| (scaled), rms(a-b) / rms(b), and north_star's per-voxel relative error
|a-b| / max(|b|, floor*max|b|) for floor = 1e-3 (SURVEY hard part 7) -- its maximum and its 99.99th percentile.
Columns: bf16 CUDA path, bf16 path with an fp32 modulator tail (tail_dtype), fp32 CUDA path, stock torch.autocast(bfloat16) of the oracle.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coma_unet_b200 as cu            # noqa: E402
from coma_unet_b200 import ops          # noqa: E402
from oracle import model as omodel     # noqa: E402
from tests.golden import common        # noqa: E402


def to_ncdhw(t, channels):
    if isinstance(t, ops.Deferred):
        t = t.materialize()
    if isinstance(t, (tuple, list)):
        t = t[0]
    if not torch.is_tensor(t):
        return None
    if t.dim() == 5:
        t = t.permute(0, 4, 1, 2, 3)
        if channels is not None and t.shape[1] > channels:
            t = t[:, :channels]
    return t.float()


def capture(model, ours, run):
    acts, order, hooks = {}, [], []
    for name, mod in model.named_modules():
        if not name:
            continue

        def hook(m, inp, out, name=name):
            ch = getattr(m, "out_channels", None) if ours else None
            t = to_ncdhw(out, ch) if ours else (out[0] if isinstance(out, (tuple, list)) else out)
            if torch.is_tensor(t) and t.dim() == 5 and name not in acts:
                acts[name] = t.detach().float()
                order.append(name)
        hooks.append(mod.register_forward_hook(hook))
    with torch.no_grad():
        pred = run(model)
    for h in hooks:
        h.remove()
    acts["<prediction>"] = pred.detach().float()
    order.append("<prediction>")
    return acts, order


def errors(a, b, floor=1e-3):
    d = (a - b).abs()
    mx = b.abs().max().clamp_min(1e-30)
    rel = d / torch.maximum(b.abs(), floor * mx)
    flat = rel.flatten()
    k = max(1, int(flat.numel() * 1e-4))
    p9999 = float(torch.topk(flat, k).values[-1]) if flat.numel() > k else float(flat.max())
    return {"scaled_max": float(d.max() / mx), "rms_rel": float(d.pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)),
            "rel_floor1e-3_max": float(flat.max()), "rel_floor1e-3_p99.99": p9999,
            "frac_voxels_rel_gt_1e-2": float((flat > 1e-2).float().mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs="+", default=[128])
    ap.add_argument("--width", type=int, default=32)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--train", action="store_true", help="train-mode forward (BatchNorm batch statistics) instead of eval")
    ap.add_argument("--out", default="gpurun_out/error_trace.json")
    args = ap.parse_args()
    shape = tuple(args.shape * 3 if len(args.shape) == 1 else args.shape)
    channels = [args.width * 2 ** i for i in range(5)]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    kw = dict(latent_spaces=[2048] * 5, conditional=True, prompt_shape=shape)
    mri, tau, roi, covars, dicts = common.synthetic_batch(args.batch, shape, args.seed)
    mri, roi = mri.to(dev), roi.to(dev)

    def run(m):
        m.train(args.train)
        m.set_training(args.train)
        out = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
        return out[0] if isinstance(out, (tuple, list)) else out

    oracle = common.fill_deterministic(omodel.ContrastiveAttentionUNET_DP(3, 1, 1, channels, [2] * 5, **kw), args.seed).to(dev)
    ref, order = capture(oracle, False, run)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        auto, _ = capture(oracle, False, run)
    del oracle
    variants = {}
    for label, mk in (("bf16", dict(compute_dtype=torch.bfloat16)),
                      ("bf16_tail_fp32", dict(compute_dtype=torch.bfloat16, tail_dtype=torch.float32)),
                      ("fp32", dict(compute_dtype=torch.float32))):
        m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, channels, [2] * 5, **kw, **mk)
        m.set_save_attn(None)
        common.fill_deterministic(m, args.seed).to(dev)
        variants[label], _ = capture(m, True, run)
        del m
        torch.cuda.empty_cache()
    variants["oracle_autocast_bf16"] = auto
    rows = []
    for name in order:
        b = ref[name]
        row = {"module": name, "shape": list(b.shape)}
        for label, acts in variants.items():
            a = acts.get(name)
            if a is not None and a.shape == b.shape:
                row[label] = errors(a, b)
        if len(row) > 2:
            rows.append(row)
    print(f"{'module':58s} " + " ".join(f"{l[:20]:>20s}" for l in variants) + "   (scaled max | rel@1e-3 p99.99)")
    for r in rows:
        cells = []
        for label in variants:
            e = r.get(label)
            cells.append(f"{e['scaled_max']:9.2e}|{e['rel_floor1e-3_p99.99']:9.2e}" if e else " " * 19)
        print(f"{r['module'][:58]:58s} " + "  ".join(cells))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"shape": list(shape), "channels": channels, "seed": args.seed, "batch": args.batch, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
