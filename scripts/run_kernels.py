"""One launch of every hot kernel family on its benchmark-size problem (for one `ncu --set full` capture and for CUDA-event times):
forward convs, their data gradients, the tcgen05 weight gradients (stride 1 and stride 2), the norm backward sweeps, the norm
apply and the fused gate.

    python scripts/run_kernels.py [reps]
    ncu --set full --clock-control none --import-source on -k regex:'conv_pair|conv_halo|convT_halo|conv_tc_kernel|wgrad_tc|wgrad_s2|bwd_bulk|affine_act_vec|gate_mma' \
        -o gpurun_out/r02_kernels python scripts/run_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_unet_b200 import _lib as L
from coma_unet_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = "cuda"


def timed(label, fn, flops=0.0, nbytes=0.0):
    fn()
    torch.cuda.synchronize()
    if reps == 0:          # under ncu: one launch per kernel is enough
        print(label, flush=True)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    extra = f"{flops / ms / 1e9:8.1f} TFLOP/s" if flops else f"{nbytes / ms / 1e6:8.0f} GB/s"
    print(f"{label:64s} {ms:7.3f} ms {extra}", flush=True)


def conv(label, B, Cin, Cout, D, k, s, T, kind=None):
    x = torch.randn(B, D, D, D, Cin, device=dev).bfloat16()
    w = torch.randn(Cin, Cout, k, k, k, device=dev) if T else torch.randn(Cout, Cin, k, k, k, device=dev)
    wp = ops.pack_weight(w, bool(T), Cin, Cout, torch.bfloat16)
    vox = D ** 3 if T else (D // s) ** 3
    timed(f"{label} B{B} {Cin}->{Cout} @{D}^3 k{k} s{s}{' T' if T else ''}",
          lambda: ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=bool(T), kind=kind), 2.0 * B * vox * k ** 3 * Cin * Cout)


def wgrad(label, B, Cg, Cx, Dg, s):
    g = torch.randn(B, Dg, Dg, Dg, Cg, device=dev).bfloat16()
    x = torch.randn(B, Dg * s, Dg * s, Dg * s, Cx, device=dev).bfloat16()
    timed(f"{label} B{B} Cg{Cg} Cx{Cx} coarse {Dg}^3 s{s}", lambda: ops.wgrad_raw(g, x, ksize=3, stride=s), 2.0 * 27 * Cg * Cx * B * Dg ** 3)


# forward (batch 8, the inference bench) and data gradients (batch 4, the training bench)
conv("fprop pair", 8, 64, 32, 128, 3, 1, 0)
conv("fprop halo3", 8, 32, 32, 128, 3, 1, 0)
conv("fprop halo3-16", 8, 16, 16, 128, 3, 1, 0)
conv("fprop s2", 8, 32, 64, 128, 3, 2, 0)
conv("fprop convT", 8, 64, 32, 64, 3, 2, 1)
conv("fprop pair 128", 8, 128, 64, 64, 3, 1, 0)
conv("fprop per-tap 256", 8, 256, 256, 16, 3, 1, 0)
conv("dgrad pair (of 64->32)", 4, 32, 64, 128, 3, 1, 0, kind="coma_conv3d_dgrad")
conv("dgrad of s2 conv (transposed)", 4, 64, 32, 64, 3, 2, 1, kind="coma_conv3d_dgrad")
conv("dgrad of convT (strided)", 4, 32, 64, 128, 3, 2, 0, kind="coma_convT3d_dgrad")
# weight gradients (batch 4)
wgrad("wgrad tc", 4, 32, 64, 128, 1)
wgrad("wgrad tc", 4, 32, 32, 128, 1)
wgrad("wgrad tc 16", 4, 16, 16, 128, 1)
wgrad("wgrad s2 tc", 4, 64, 32, 64, 2)
wgrad("wgrad s2 tc", 4, 128, 64, 32, 2)
# norm backward (reduce + finalize + apply), norm apply, gate
for C, D in ((32, 128), (64, 64)):
    x = torch.randn(4, D, D, D, C, device=dev).bfloat16().requires_grad_(True)
    gpar = (torch.rand(4, C, device=dev) + 0.5).requires_grad_(True)
    hpar = torch.randn(4, C, device=dev).requires_grad_(True)
    dy = torch.randn(4, D, D, D, C, device=dev).bfloat16()
    y = ops.norm_act(x, gpar, hpar, None, ops.NormCfg(mode=L.NORM_BATCH, act=L.ACT_RELU))
    timed(f"norm backward B4 C{C} @{D}^3 (5 tensor passes)", lambda: torch.autograd.grad(y, (x, gpar, hpar), dy, retain_graph=True), nbytes=5.0 * x.numel() * 2)
    A, S = torch.rand(4, C, device=dev) + 0.5, torch.randn(4, C, device=dev)
    out = torch.empty_like(x)
    timed(f"norm apply B4 C{C} @{D}^3 (2 tensor passes)", lambda: ops.affine_act(x.detach(), A, S, None, L.ACT_RELU, out=out), nbytes=2.0 * x.numel() * 2)
C, D, B = 32, 128, 8
g, x = torch.randn(B, D, D, D, C, device=dev).bfloat16(), torch.randn(B, D, D, D, C, device=dev).bfloat16()
wg, wx = torch.randn(C // 2, C, device=dev) * 0.1, torch.randn(C // 2, C, device=dev) * 0.1
bsum, wpsi, bpsi = torch.zeros(C // 2, device=dev), torch.randn(C // 2, device=dev), torch.zeros(1, device=dev)
o = torch.empty_like(x)
timed(f"gate fused B{B} C{C} @{D}^3 (3 tensor passes)", lambda: ops.gate_fused(g, x, wg, wx, bsum, wpsi, bpsi, out=o), nbytes=3.0 * x.numel() * 2)
