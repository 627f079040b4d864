"""GPU: every kernel family through the C ABI against plain fp32 torch references on seeded inputs.

Tolerances: fp32 storage <= 1e-4 (north_star's TF32/fp32 bound), bf16 storage <= 1e-2, both as
max |a-b| / max|b| unless noted.  Edge cases: ragged sizes, Cin = 1, per-sample weights, channel-sliced
(concat-buffer) views, padded channels, B = 1.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from coma_unet_b200 import _lib as L   # noqa: E402
from coma_unet_b200 import ops          # noqa: E402

DEV = "cuda"
TOL = {torch.float32: 1e-4, torch.bfloat16: 1.2e-2}


@pytest.fixture(autouse=True)
def _exact_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(*shape, generator=g)).to(DEV)


def to_vol(x, dtype):   # NCDHW fp32 -> NDHWC dtype
    return x.permute(0, 2, 3, 4, 1).contiguous().to(dtype)


def to_ncdhw(v):
    return v.permute(0, 4, 1, 2, 3).float()


CONV_CASES = [
    # (B, Cin, Cout, D, H, W, k, stride, transposed)
    (2, 16, 32, 8, 8, 8, 3, 1, False),
    (1, 32, 32, 9, 7, 10, 3, 1, False),     # ragged
    (2, 1, 16, 8, 8, 8, 3, 1, False),       # Cin = 1 (head conv)
    (2, 32, 64, 8, 8, 8, 3, 2, False),
    (1, 16, 16, 6, 10, 12, 3, 2, False),
    (2, 64, 32, 4, 4, 4, 3, 2, True),
    (1, 32, 16, 3, 5, 4, 3, 2, True),
    (2, 32, 16, 8, 8, 8, 1, 1, False),
    (2, 16, 1, 8, 8, 8, 1, 1, False),       # psi / head convs
    (2, 32, 1, 17, 18, 20, 1, 1, False),    # one-output pointwise weight gradient: streaming kernel, several blocks, ragged tail
    (1, 64, 1, 16, 16, 16, 1, 1, False),
    (3, 8, 1, 9, 10, 12, 1, 1, False),
    (2, 2, 1, 9, 10, 12, 1, 1, False),      # final_pred_head 2 -> 1: dense few-channel x, 4 voxels per 16-byte load, ragged tail
    (1, 4, 1, 33, 32, 40, 1, 1, False),     # several blocks, the four-deep main loop
    (2, 1, 1, 16, 16, 24, 1, 1, False),
    (1, 16, 16, 9, 10, 12, 1, 1, False),    # few-channel pointwise streaming kernel, ragged last chunk
    (2, 32, 32, 16, 16, 16, 1, 1, False),   # two statistics chunks per sample
    (2, 64, 32, 9, 10, 12, 1, 1, False),    # second-level gate convs (64 -> 32) and their data gradient (32 -> 64) on the streaming kernel
    (1, 32, 64, 16, 16, 16, 1, 1, False),
    (1, 3, 5, 5, 6, 7, 3, 1, False),        # odd channel counts (scalar path)
    (2, 16, 1, 9, 10, 12, 3, 1, False),     # modulator head 16 -> 1: single-pass one-channel weight gradient
    (1, 32, 1, 8, 8, 8, 3, 1, False),
    # >= 16^3 voxels, k3 s1, <= 64 channels: halo weight-gradient kernel (bf16) / halo forward + dgrad kernels
    (2, 16, 32, 16, 16, 16, 3, 1, False),
    (1, 32, 32, 17, 18, 20, 3, 1, False),   # ragged in every dim
    (1, 64, 32, 16, 16, 16, 3, 1, False),
    (1, 16, 16, 12, 24, 16, 3, 1, False),
    (1, 64, 128, 16, 16, 16, 3, 1, False),  # halo wgrad tiled over 2 x 4 channel tiles
    # stride-2 halo weight gradient (down-sampling convs; transposed convs with x / dy swapped)
    (1, 32, 32, 16, 16, 16, 3, 2, False),
    (1, 16, 32, 18, 20, 22, 3, 2, False),   # ragged blocks in every dim
    (1, 32, 16, 8, 8, 8, 3, 2, True),
    (2, 64, 32, 9, 10, 12, 3, 2, True),
    # tcgen05 weight gradient (k3 s1, channels multiples of 32, W % 32 == 0, H % 8 == 0): MN-major operands, nine taps per MMA
    (1, 32, 32, 5, 8, 32, 3, 1, False),     # one column, one channel pair
    (2, 64, 32, 9, 16, 64, 3, 1, False),    # two x slabs, several columns, odd depth (ragged depth chunks)
    (1, 32, 128, 20, 24, 32, 3, 1, False),  # four gradient tiles, depth split into chunks
    (1, 16, 16, 5, 8, 32, 3, 1, False),     # 16-channel rows on both operands (32B swizzle, eight row-shifted chunks in M)
    (2, 32, 16, 5, 16, 32, 3, 1, False),    # mixed row widths
    (1, 16, 32, 4, 8, 64, 3, 1, False),
    (2, 32, 64, 16, 16, 16, 3, 1, False),   # 16-voxel lines (the 16^3 level): 16 lines x 16 voxels per plane
    (2, 16, 1, 6, 16, 32, 3, 1, False),     # one-channel gradient padded to a 16-channel row for the tcgen05 weight gradient
    # tcgen05 STRIDE-2 weight gradient (coarse grid W % 32 == 0, H % 4 == 0 or W % 16 == 0, H % 8 == 0; channels multiples of 32):
    # fine planes split into even / odd W slabs by TMA element strides, kh in the M chunks, (kd, kw) in nine accumulators
    (1, 32, 32, 6, 8, 64, 3, 2, False),     # one column, 3 coarse planes
    (2, 32, 64, 10, 16, 64, 3, 2, False),   # two gradient tiles, several H columns, odd number of coarse planes
    (1, 64, 32, 8, 16, 32, 3, 2, False),    # 16-voxel lines (coarse 4 x 8 x 16), two x slabs
    (2, 64, 64, 20, 32, 32, 3, 2, False),   # depth split into chunks
    (1, 64, 32, 5, 4, 32, 3, 2, True),      # transposed conv: x / dy swap roles (coarse = the conv-transpose INPUT)
    (2, 32, 32, 4, 8, 16, 3, 2, True),
    (1, 128, 64, 3, 8, 32, 3, 2, True),
    # one-channel k3 weight gradient on mma.sync with gathered A fragments (Cx = 16, W % 32 == 0, H % 8 == 0)
    (2, 16, 1, 5, 8, 32, 3, 1, False),
    (1, 16, 1, 20, 24, 64, 3, 1, False),    # several blocks in every dim, depth chunks
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_bwd(case, dtype):
    B, Cin, Cout, D, H, W, k, s, tr = case
    x = rnd(B, Cin, D, H, W, seed=1)
    wshape = (Cin, Cout, k, k, k) if tr else (Cout, Cin, k, k, k)
    w = rnd(*wshape, seed=2, scale=(Cin * k ** 3) ** -0.5)
    b = rnd(Cout, seed=3, scale=0.1)
    if dtype == torch.bfloat16:   # make inputs exactly representable so only accumulation order differs
        x, w = x.bfloat16().float(), w.bfloat16().float()
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pad = (k - 1) // 2
    if tr:
        ref = F.conv_transpose3d(xr, wr, br, stride=s, padding=pad, output_padding=s - 1)
    else:
        ref = F.conv3d(xr, wr, br, stride=s, padding=pad)
    gy = rnd(*ref.shape, seed=4)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    ref.backward(gy)

    xv = to_vol(x, dtype).requires_grad_(True)
    wp, bp = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y, stats = ops.conv3d(xv, wp, bp, ops.ConvCfg(ksize=k, stride=s, transposed=tr, want_stats=True))
    assert err(to_ncdhw(y), ref) < TOL[dtype]
    # fused statistics of the raw output
    tot = stats.sum(dim=1)[:, :Cout]
    assert err(tot[..., 0], ref.detach().sum(dim=(2, 3, 4))) < 5 * TOL[dtype]
    assert err(tot[..., 1], (ref.detach() ** 2).sum(dim=(2, 3, 4))) < 5 * TOL[dtype]
    y.backward(to_vol(gy, dtype))
    assert err(to_ncdhw(xv.grad), xr.grad) < TOL[dtype]
    assert err(wp.grad, wr.grad) < TOL[dtype]
    assert err(bp.grad, br.grad) < TOL[dtype]


SPLIT_CASES = [(2, 16, 32, 8, 8, 8, 3, 1, False), (1, 32, 32, 9, 7, 10, 3, 1, False), (2, 1, 16, 8, 8, 8, 3, 1, False),
               (2, 32, 64, 8, 8, 8, 3, 2, False), (2, 64, 32, 4, 4, 4, 3, 2, True), (1, 3, 5, 5, 6, 7, 3, 1, False),
               (2, 16, 1, 9, 10, 12, 3, 1, False), (1, 64, 128, 16, 16, 16, 3, 1, False), (2, 32, 32, 16, 16, 32, 3, 1, False)]


@pytest.mark.parametrize("case", SPLIT_CASES)
def test_fp32_conv_on_tensor_cores_split_precision(case):
    """ops.fp32_split: fp32 tensors, 3x3x3 convolutions on tcgen05 through bf16 hi/lo operands and an fp32-stored accumulator
    (dtype COMA_BF16_F32OUT).  Forward, data gradient, weight gradient, bias gradient and the fused statistics hold the fp32
    tolerance 1e-4 against cuDNN fp32 on arbitrary (not bf16-representable) inputs."""
    B, Cin, Cout, D, H, W, k, s, tr = case
    x = rnd(B, Cin, D, H, W, seed=1)
    wshape = (Cin, Cout, k, k, k) if tr else (Cout, Cin, k, k, k)
    w = rnd(*wshape, seed=2, scale=(Cin * k ** 3) ** -0.5)
    b = rnd(Cout, seed=3, scale=0.1)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv_transpose3d(xr, wr, br, stride=s, padding=1, output_padding=s - 1) if tr else F.conv3d(xr, wr, br, stride=s, padding=1)
    gy = rnd(*ref.shape, seed=4)
    ref.backward(gy)
    xv = to_vol(x, torch.float32).requires_grad_(True)
    wp, bp = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    l0 = L.launches
    with ops.fp32_split(True):
        y, stats = ops.conv3d(xv, wp, bp, ops.ConvCfg(ksize=k, stride=s, transposed=tr, want_stats=True))
    assert y.dtype == torch.float32 and err(to_ncdhw(y), ref) < 1e-4
    tot = stats.sum(dim=1)[:, :Cout]
    assert err(tot[..., 0], ref.detach().sum(dim=(2, 3, 4))) < 5e-4
    assert err(tot[..., 1], (ref.detach() ** 2).sum(dim=(2, 3, 4))) < 5e-4
    y.backward(to_vol(gy, torch.float32))           # the switch travels with the autograd node
    assert L.launches - l0 >= 3
    assert err(to_ncdhw(xv.grad), xr.grad) < 1e-4
    assert err(wp.grad, wr.grad) < 1e-4
    assert err(bp.grad, br.grad) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_epilogue_and_channel_sliced_io(dtype):
    B, Cin, Cout, D = 2, 32, 32, 8
    x = rnd(B, Cin, D, D, D, seed=5)
    w = rnd(Cout, Cin, 3, 3, 3, seed=6, scale=0.03)
    b = rnd(Cout, seed=7, scale=0.1)
    scale, shift = 1 + 0.2 * rnd(B, Cout, seed=8), 0.1 * rnd(B, Cout, seed=9)
    slope = torch.tensor([0.2], device=DEV)
    if dtype == torch.bfloat16:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    ref = F.conv3d(x, w, b, padding=1) * scale[:, :, None, None, None] + shift[:, :, None, None, None]
    ref = torch.where(ref > 0, ref, 0.2 * ref)
    xbuf = torch.zeros(B, D, D, D, 2 * Cin, device=DEV, dtype=dtype)
    xbuf[..., Cin:] = to_vol(x, dtype)
    ybuf = torch.full((B, D, D, D, 2 * Cout), 7.0, device=DEV, dtype=dtype)
    wp = ops.pack_weight(w, False, Cin, Cout, dtype)
    ops.conv_raw(xbuf[..., Cin:], wp, b, ksize=3, scale=scale.contiguous(), shift=shift.contiguous(), slope=slope,
                 act=L.ACT_LEAKY, out=ybuf[..., :Cout])
    assert err(to_ncdhw(ybuf[..., :Cout]), ref) < TOL[dtype]
    assert bool((ybuf[..., Cout:] == 7.0).all())        # the other half of the concat buffer is untouched


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_per_sample_conv1x1(dtype):
    B, Cin, Cout, D = 3, 32, 1, 6
    x = rnd(B, Cin, D, D, D, seed=10)
    w = rnd(B, Cout, Cin, seed=11, scale=0.2)
    b = rnd(B, Cout, seed=12, scale=0.1)
    if dtype == torch.bfloat16:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.einsum("bcdhw,boc->bodhw", xr, wr) + br[:, :, None, None, None]
    gy = rnd(*ref.shape, seed=13)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    ref.backward(gy)
    xv = to_vol(x, dtype).requires_grad_(True)
    wp, bp = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = ops.PerSampleConv1x1Fn.apply(xv, wp, bp)
    assert err(to_ncdhw(y), ref) < TOL[dtype]
    y.backward(to_vol(gy, dtype))
    assert err(to_ncdhw(xv.grad), xr.grad) < TOL[dtype]
    assert err(wp.grad, wr.grad) < TOL[dtype]
    assert err(bp.grad, br.grad) < TOL[dtype]


NORM_CASES = [
    # (mode, act, C, film, residual)
    (L.NORM_INSTANCE, L.ACT_LEAKY, 32, True, False),
    (L.NORM_INSTANCE, L.ACT_LEAKY, 1, False, False),
    (L.NORM_INSTANCE, L.ACT_LEAKY_RELU, 1, False, False),
    (L.NORM_BATCH, L.ACT_RELU, 64, True, False),
    (L.NORM_BATCH, L.ACT_NONE, 16, False, False),
    (L.NORM_BATCH, L.ACT_RELU, 16, False, True),
    (L.NORM_BATCH, L.ACT_SIGMOID, 1, False, False),
    (L.NORM_GIVEN, L.ACT_RELU, 32, True, False),
    (L.NORM_NONE, L.ACT_LEAKY, 8, False, False),
]


def norm_reference(x, g, h, slope, mode, act, rm, rv, residual, eps=1e-5):
    if mode == L.NORM_INSTANCE:
        mean, var = x.mean(dim=(2, 3, 4), keepdim=True), x.var(dim=(2, 3, 4), unbiased=False, keepdim=True)
    elif mode == L.NORM_BATCH:
        mean, var = x.mean(dim=(0, 2, 3, 4), keepdim=True), x.var(dim=(0, 2, 3, 4), unbiased=False, keepdim=True)
    elif mode == L.NORM_GIVEN:
        mean, var = rm[None, :, None, None, None], rv[None, :, None, None, None]
    else:
        mean, var = torch.zeros_like(x[:1, :, :1, :1, :1]), torch.ones_like(x[:1, :, :1, :1, :1]) - eps
    xh = (x - mean) / torch.sqrt(var + eps)
    u = xh * g[:, :, None, None, None] + h[:, :, None, None, None]
    if residual is not None:
        u = u + residual
    if act == L.ACT_RELU:
        return torch.relu(u)
    if act == L.ACT_LEAKY:
        return torch.where(u > 0, u, slope * u)
    if act == L.ACT_LEAKY_RELU:
        return torch.relu(torch.where(u > 0, u, slope * u))
    if act == L.ACT_SIGMOID:
        return torch.sigmoid(u)
    return u


# shapes big enough for the bulk-copy streaming sweeps of the backward (bf16, contiguous, V*C >= 32768): whole tiles, a ragged
# last tile, more chunks than tiles, one 16-voxel tile per 512-channel row block
NORM_BULK_CASES = [
    (L.NORM_BATCH, L.ACT_RELU, 32, True, False, (2, 16, 16, 16)),
    (L.NORM_INSTANCE, L.ACT_LEAKY, 16, True, False, (2, 17, 18, 20)),
    (L.NORM_BATCH, L.ACT_RELU, 16, False, True, (2, 12, 16, 24)),
    (L.NORM_INSTANCE, L.ACT_LEAKY_RELU, 8, False, False, (1, 20, 24, 16)),
    (L.NORM_BATCH, L.ACT_RELU, 256, True, False, (2, 8, 8, 8)),
    (L.NORM_BATCH, L.ACT_RELU, 512, False, False, (3, 4, 4, 4)),
    (L.NORM_BATCH, L.ACT_SIGMOID, 64, False, True, (1, 9, 10, 12)),
    (L.NORM_INSTANCE, L.ACT_NONE, 128, False, False, (2, 16, 16, 16)),
    # dense one-channel volumes (gate psi, modulator heads): the 8-voxels-per-load sweeps, and a ragged size that must not take them
    (L.NORM_INSTANCE, L.ACT_LEAKY_RELU, 1, False, False, (2, 32, 32, 32)),
    (L.NORM_BATCH, L.ACT_SIGMOID, 1, False, False, (2, 16, 24, 40)),
    (L.NORM_BATCH, L.ACT_RELU, 1, False, True, (1, 16, 16, 16)),
    (L.NORM_INSTANCE, L.ACT_LEAKY, 1, False, False, (2, 7, 9, 11)),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", NORM_BULK_CASES)
def test_norm_film_act_fwd_bwd_streaming_sizes(case, dtype):
    test_norm_film_act_fwd_bwd(case[:5], dtype, shape=case[5])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", NORM_CASES)
def test_norm_film_act_fwd_bwd(case, dtype, shape=(2, 6, 5, 8)):
    mode, act, Cn, film, use_res = case
    B, D, H, W = shape
    x = rnd(B, Cn, D, H, W, seed=20) * 1.5 + 0.3
    res = rnd(B, Cn, D, H, W, seed=26) if use_res else None
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
        res = None if res is None else res.bfloat16().float()
    g = (1 + 0.3 * rnd(B, Cn, seed=21)) if film else (1 + 0.3 * rnd(1, Cn, seed=21)).expand(B, Cn).contiguous()
    h = 0.2 * rnd(B, Cn, seed=22) if film else (0.2 * rnd(1, Cn, seed=22)).expand(B, Cn).contiguous()
    slope0 = -0.15 if act == L.ACT_LEAKY_RELU else 0.25
    rm, rv = 0.1 * rnd(Cn, seed=23), 1 + 0.2 * rnd(Cn, seed=24).abs()
    xr, gr, hr = x.clone().requires_grad_(True), g.clone().requires_grad_(True), h.clone().requires_grad_(True)
    sr = torch.tensor([slope0], device=DEV, requires_grad=True)
    rr = None if res is None else res.clone().requires_grad_(True)
    ref = norm_reference(xr, gr, hr, sr, mode, act, rm, rv, rr)
    gy = rnd(*ref.shape, seed=25)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    ref.backward(gy)

    xv = to_vol(x, dtype).requires_grad_(True)
    gp, hp = g.clone().requires_grad_(True), h.clone().requires_grad_(True)
    sp = torch.tensor([slope0], device=DEV, requires_grad=True)
    rp = None if res is None else to_vol(res, dtype).requires_grad_(True)
    run_m, run_v = rm.clone(), rv.clone()
    cfg = ops.NormCfg(mode=mode, act=act, running_mean=run_m, running_var=run_v, update_running=mode == L.NORM_BATCH,
                      n_updates=2)
    has_slope = act in (L.ACT_LEAKY, L.ACT_LEAKY_RELU)
    y = ops.norm_act(xv, gp, hp, sp if has_slope else None, cfg, residual=rp)
    tol = TOL[dtype]
    assert err(to_ncdhw(y), ref) < tol
    y.backward(to_vol(gy, dtype))
    btol = 3 * tol
    assert err(to_ncdhw(xv.grad), xr.grad) < btol
    assert err(gp.grad, gr.grad) < btol
    assert err(hp.grad, hr.grad) < btol
    if has_slope:
        assert err(sp.grad, sr.grad) < btol
    if rp is not None:
        assert err(to_ncdhw(rp.grad), rr.grad) < btol
    if mode == L.NORM_BATCH:   # two momentum updates with the same batch statistics (duplicated reference forward)
        n = B * D * H * W
        bm, bv = x.mean(dim=(0, 2, 3, 4)), x.var(dim=(0, 2, 3, 4), unbiased=False) * n / (n - 1)
        em, ev = rm.clone(), rv.clone()
        for _ in range(2):
            em, ev = 0.9 * em + 0.1 * bm, 0.9 * ev + 0.1 * bv
        assert err(run_m, em) < 1e-4 and err(run_v, ev) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Cn", [16, 32, 64])
def test_fused_gate_matches_block(Cn, dtype):
    B, D = 2, 8
    Fi = Cn // 2
    g, x = rnd(B, Cn, D, D, D, seed=30), rnd(B, Cn, D, D, D, seed=31)
    if dtype == torch.bfloat16:
        g, x = g.bfloat16().float(), x.bfloat16().float()
    wg, wx = rnd(Fi, Cn, seed=32, scale=Cn ** -0.5), rnd(Fi, Cn, seed=33, scale=Cn ** -0.5)
    bsum, wpsi, bpsi = 0.1 * rnd(Fi, seed=34), rnd(Fi, seed=35, scale=Fi ** -0.5), 0.1 * rnd(1, seed=36)
    t = torch.relu(torch.einsum("bcdhw,fc->bfdhw", g, wg) + torch.einsum("bcdhw,fc->bfdhw", x, wx) + bsum[None, :, None, None, None])
    psi = torch.sigmoid(torch.einsum("bfdhw,f->bdhw", t, wpsi) + bpsi)[:, None]
    ref = x * psi
    cat = torch.zeros(B, D, D, D, 2 * Cn, device=DEV, dtype=dtype)
    cat[..., Cn:] = to_vol(g, dtype)
    coeff = torch.empty(B, D, D, D, 1, device=DEV, dtype=dtype)
    ops.gate_fused(cat[..., Cn:], to_vol(x, dtype), wg, wx, bsum, wpsi, bpsi, out=cat[..., :Cn], psi_out=coeff)
    assert err(to_ncdhw(cat[..., :Cn]), ref) < TOL[dtype]
    assert err(to_ncdhw(coeff), psi) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bcast_mul_and_concat(dtype):
    B, Cn, D = 2, 32, 6
    x, p = rnd(B, Cn, D, D, D, seed=40), torch.sigmoid(rnd(B, 1, D, D, D, seed=41))
    if dtype == torch.bfloat16:
        x, p = x.bfloat16().float(), p.bfloat16().float()
    xr, pr = x.clone().requires_grad_(True), p.clone().requires_grad_(True)
    ref = xr * pr
    gy = rnd(*ref.shape, seed=42)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    ref.backward(gy)
    xv, pv = to_vol(x, dtype).requires_grad_(True), to_vol(p, dtype).requires_grad_(True)
    y = ops.bcast_mul(xv, pv)
    assert err(to_ncdhw(y), ref) < TOL[dtype]
    y.backward(to_vol(gy, dtype))
    assert err(to_ncdhw(xv.grad), xr.grad) < TOL[dtype]
    assert err(to_ncdhw(pv.grad), pr.grad) < 2 * TOL[dtype]
    a, b = to_vol(x, dtype).requires_grad_(True), to_vol(gy, dtype).requires_grad_(True)
    c = ops.concat2(a, b)
    assert torch.equal(c, torch.cat((a, b), dim=-1))
    c.backward(torch.cat((b, a), dim=-1).detach())
    assert torch.equal(a.grad, b.detach()) and torch.equal(b.grad, a.detach())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_roi_paint_pack_and_mse(dtype):
    from tests.golden import common
    B, shape = 3, (12, 10, 16)
    mri, tau, roi, covars, dicts = common.synthetic_batch(B, shape, 9)
    mri, tau, roi = mri.to(DEV), tau.to(DEV), roi.to(DEV)
    names = common.roi_names()
    V = shape[0] * shape[1] * shape[2]
    lut = torch.tensor([[[d[n]["loc"], d[n]["std"]] for n in names] for d in dicts], device=DEV)
    ids = torch.tensor(common.ROI_INDICES, dtype=torch.int32, device=DEV)
    is_pos = (covars[:, 0, 0] == 1).float().to(DEV)
    pos, neg = rnd(1, 1, *shape, seed=50).requires_grad_(True), rnd(1, 1, *shape, seed=51).requires_grad_(True)
    out = ops.RoiPaintFn.apply(pos, neg, roi, mri, lut, ids, is_pos, 16, dtype, (True, True))
    suvr, sal = torch.zeros_like(mri), torch.zeros_like(mri)
    for b in range(B):
        for i, idx in enumerate(common.ROI_INDICES):
            suvr[b][roi[b] == idx] = lut[b, i, 0]
            sal[b][roi[b] == idx] = lut[b, i, 1]
    suvr, sal = torch.where(mri < 1e-4, 0 * suvr, suvr), torch.where(mri < 1e-4, 0 * sal, sal)
    prompt = torch.where(is_pos.view(B, 1, 1, 1, 1) == 1, pos.detach(), neg.detach())
    assert err(to_ncdhw(out[..., 0:1]), prompt) < TOL[dtype]
    assert err(to_ncdhw(out[..., 1:2]), sal) < TOL[dtype] and err(to_ncdhw(out[..., 2:3]), suvr) < TOL[dtype]
    assert float(out[..., 3:].abs().max()) == 0.0
    gbuf = rnd(B, *shape, 16, seed=52).to(dtype)
    out.backward(gbuf)
    g0 = gbuf[..., 0].float()
    assert err(pos.grad.reshape(-1), (g0 * is_pos.view(B, 1, 1, 1)).sum(0).reshape(-1)) < TOL[dtype]
    assert err(neg.grad.reshape(-1), (g0 * (1 - is_pos).view(B, 1, 1, 1)).sum(0).reshape(-1)) < TOL[dtype]

    a, bb, add = rnd(B, *shape, 1, seed=53).to(dtype).requires_grad_(True), rnd(B, *shape, 1, seed=54).to(dtype).requires_grad_(True), rnd(1, 1, *shape, seed=55).requires_grad_(True)
    packed = ops.Pack2Fn.apply(a, add, bb, 16)
    assert err(packed[..., 0].float(), a.detach().float()[..., 0] + add.detach().reshape(1, *shape)) < TOL[dtype]
    assert torch.equal(packed[..., 1], bb.detach()[..., 0]) and float(packed[..., 2:].abs().max()) == 0.0
    packed.backward(gbuf)
    assert torch.equal(a.grad[..., 0], gbuf[..., 0]) and torch.equal(bb.grad[..., 0], gbuf[..., 1])
    assert err(add.grad.reshape(-1), gbuf[..., 0].float().sum(0).reshape(-1)) < TOL[dtype]

    from oracle import criterions as ocrit
    import coma_unet_b200 as cu
    w = torch.tensor([225.0] * 36)
    pred = (tau + 0.2 * rnd(*tau.shape, seed=56)).to(dtype).float()
    pr, pp = pred.clone().requires_grad_(True), pred.clone().to(dtype).requires_grad_(True)
    o_loss, p_loss = ocrit.RoiMSE(w, common.ROI_INDICES, reduction=None, voxel_wise=False), cu.RoiMSE(w, common.ROI_INDICES, reduction=None, voxel_wise=False)
    lr, lp = o_loss(pr, tau, roi), p_loss(pp, tau, roi)
    assert lp.shape == lr.shape and err(lp, lr) < 1e-4
    coef = rnd(B, 1, seed=57)
    (lr * coef).sum().backward()
    (lp * coef).sum().backward()
    assert err(pp.grad.float(), pr.grad) < TOL[dtype]
    assert err(cu.RoiMSE(w, common.ROI_INDICES, voxel_wise=False)(pp.detach(), tau, roi), ocrit.RoiMSE(w, common.ROI_INDICES, voxel_wise=False)(pr.detach(), tau, roi)) < 1e-4


TC_CASES = [
    # (B, Cin, Cout, D, H, W, k, stride, transposed)
    (2, 32, 32, 16, 16, 16, 3, 1, False),
    (1, 64, 32, 12, 9, 20, 3, 1, False),     # ragged tiles, K chunk 64 (128B swizzle)
    (1, 16, 16, 8, 8, 8, 3, 1, False),       # 32B swizzle
    (2, 128, 64, 8, 8, 8, 3, 1, False),      # two K chunks per tap
    (1, 64, 512, 4, 4, 8, 3, 1, False),      # two N tiles
    (2, 32, 64, 16, 16, 16, 3, 2, False),    # TMA element stride 2
    (1, 64, 128, 10, 6, 12, 3, 2, False),
    (2, 64, 32, 8, 8, 8, 3, 2, True),        # transposed: 8 parity classes
    (1, 32, 16, 5, 3, 6, 3, 2, True),
    (2, 64, 32, 8, 8, 8, 1, 1, False),
    # halo-reuse kernel (k3 s1, Cin/Cout <= 64, planes >= 16 x 8)
    (1, 64, 32, 6, 20, 13, 3, 1, False),     # ragged in every dim, two d-segments
    (1, 16, 16, 9, 16, 8, 3, 1, False),      # exactly one tile per plane, 32B swizzle
    (2, 32, 64, 5, 32, 24, 3, 1, False),
    (1, 32, 16, 20, 17, 9, 3, 1, False),
    (1, 64, 64, 6, 16, 16, 3, 1, False),     # weights do not fit: Cout split over grid.y (2 x 32)
    (1, 64, 128, 4, 16, 8, 3, 1, False),     # 4 x 32
    (2, 32, 128, 3, 32, 8, 3, 1, False),     # 2 x 64
    (1, 128, 64, 5, 16, 16, 3, 1, False),    # Cin = 128: two 64-channel K chunks per slab, Cout split 4 x 16
    (1, 128, 32, 3, 18, 9, 3, 1, False),
    # halo-reuse transposed kernel (8 parity classes in 8 TMEM accumulators)
    (2, 64, 32, 6, 16, 8, 3, 2, True),
    (1, 32, 16, 5, 20, 13, 3, 2, True),      # ragged
    (1, 64, 64, 4, 16, 16, 3, 2, True),      # single accumulator set (8 x 64 columns)
    (1, 16, 16, 9, 17, 9, 3, 2, True),
    # stride-2 plane-ring kernel (parity sub-slabs; output planes >= 16 x 8)
    (2, 32, 64, 16, 32, 16, 3, 2, False),    # one tile per output plane
    (1, 64, 32, 11, 37, 19, 3, 2, False),    # odd input sizes, ragged tiles
    (1, 16, 16, 20, 32, 34, 3, 2, False),    # 32B swizzle, two epilogue warpgroups
    (1, 32, 128, 9, 32, 16, 3, 2, False),    # Cout split 2 x 64
    (3, 32, 32, 40, 64, 32, 3, 2, False),    # many segments per CTA: ring wrap and phase bookkeeping
    # CTA-pair kernel (cta_group::2, Cin = 128, even number of 8-wide columns)
    (2, 128, 64, 13, 32, 32, 3, 1, False),   # two Cout splits, several units per pair, two depth segments
    (1, 128, 32, 27, 17, 16, 3, 1, False),   # ragged h, three depth segments
    (1, 128, 128, 12, 16, 16, 3, 1, False),  # four Cout splits, a full 12-plane segment
    (2, 64, 64, 9, 20, 40, 3, 1, False),     # one K chunk, ragged h, odd number of units per pair
    # tap-packed kernel for 1 / 2 / 4 input channels (K = 27 * Cin, A tiles gathered by builder warps)
    (2, 1, 32, 12, 32, 16, 3, 1, False),     # the head conv on the 1-channel MRI
    (1, 2, 16, 9, 17, 12, 3, 1, False),      # ragged tiles, two K chunks (TMA needs W * Cin to be a multiple of 8)
    (1, 4, 16, 20, 32, 24, 3, 1, False),     # four K chunks (the 3-channel painted prompt + one zero channel)
    (3, 1, 16, 37, 16, 8, 3, 1, False),      # one tile per plane, long segments: stage and ring wrap
]


@pytest.mark.parametrize("case", TC_CASES)
def test_tcgen05_conv_matches_cuda_core_conv(case):
    """The tensor-core implicit GEMM, forced, against the exact direct conv on the same bf16 data (and torch)."""
    B, Cin, Cout, D, H, W, k, s, tr = case
    dtype = torch.bfloat16
    x = rnd(B, Cin, D, H, W, seed=60).bfloat16().float()
    wshape = (Cin, Cout, k, k, k) if tr else (Cout, Cin, k, k, k)
    w = rnd(*wshape, seed=61, scale=(Cin * k ** 3) ** -0.5).bfloat16().float()
    b = rnd(Cout, seed=62, scale=0.1)
    scale, shift = (1 + 0.2 * rnd(B, Cout, seed=63)).contiguous(), (0.1 * rnd(B, Cout, seed=64)).contiguous()
    pad = (k - 1) // 2
    ref = F.conv_transpose3d(x, w, b, stride=s, padding=pad, output_padding=s - 1) if tr else F.conv3d(x, w, b, stride=s, padding=pad)
    wp = ops.pack_weight(w, tr, Cin, Cout, dtype)
    xv = to_vol(x, dtype)
    outs = {}
    for impl in (L.IMPL_TCGEN05, L.IMPL_SIMT):
        y, st = ops.conv_raw(xv, wp, b, ksize=k, stride=s, transposed=tr, want_stats=True, impl=impl)
        ye, ste = ops.conv_raw(xv, wp, b, ksize=k, stride=s, transposed=tr, scale=scale, shift=shift, act=L.ACT_RELU,
                               want_stats=True, impl=impl)
        assert err(ste.sum(dim=1), st.sum(dim=1)) < 1e-5      # statistics are those of conv + bias, whatever the epilogue
        outs[impl] = (y.float(), st.sum(dim=1), ye.float())
    yt, st_t, yet = outs[L.IMPL_TCGEN05]
    ys, st_s, yes = outs[L.IMPL_SIMT]
    assert err(to_ncdhw(yt), ref) < 1e-2
    assert err(yt, ys) < 1e-2 and err(yet, yes) < 1e-2
    assert err(st_t, st_s) < 1e-3
    refe = torch.relu(ref * scale[:, :, None, None, None] + shift[:, :, None, None, None])
    assert err(to_ncdhw(yet), refe) < 1e-2


def test_tcgen05_is_selected_for_hot_path_shapes():
    import ctypes
    x = torch.zeros(1, 8, 8, 8, 32, device=DEV, dtype=torch.bfloat16)
    wp = torch.zeros(27, 32, 32, device=DEV, dtype=torch.bfloat16)
    y = torch.empty(1, 8, 8, 8, 32, device=DEV, dtype=torch.bfloat16)
    a = ops._conv_args(x, wp, None, y, ksize=3, stride=1, transposed=False, cout_comp=32)
    assert L.lib().coma_conv3d_tcgen05_supported(ctypes.byref(a)) == 1
    a32 = ops._conv_args(x.float(), wp.float(), None, y.float(), ksize=3, stride=1, transposed=False, cout_comp=32)
    assert L.lib().coma_conv3d_tcgen05_supported(ctypes.byref(a32)) == 0


# ---- input prologue: a consumer conv applies its producer's norm + FiLM + activation on load ----------------------------
PROLOGUE_CASES = [
    # (B, Cin, Cout, D, H, W, k, in_act, fused path expected)
    (2, 16, 16, 20, 32, 24, 3, "leaky", True),      # modulator stack shape: v3 plane-ring kernel, 3 CTAs / SM, 32B swizzle
    (1, 32, 32, 9, 17, 13, 3, "relu", True),        # ragged tiles, 64B swizzle, 2 CTAs / SM
    (3, 16, 32, 6, 16, 8, 3, "none", True),         # exactly one tile per plane, several samples (coefficient restaging)
    (2, 32, 1, 16, 16, 16, 1, "leaky", True),       # reduce_channels: pointwise, one output channel
    (1, 64, 16, 8, 16, 8, 3, "leaky", False),       # no fused kernel for this shape: materialised, same numbers
    (1, 16, 16, 6, 6, 6, 3, "relu", False),         # planes too small for the halo kernel
]


@pytest.mark.parametrize("case", PROLOGUE_CASES)
def test_conv_input_prologue_matches_materialised_input(case):
    import ctypes
    B, Cin, Cout, D, H, W, k, act_name, fused = case
    act = {"none": L.ACT_NONE, "relu": L.ACT_RELU, "leaky": L.ACT_LEAKY}[act_name]
    dtype = torch.bfloat16
    x = rnd(B, Cin, D, H, W, seed=90).bfloat16().float()
    w = rnd(Cout, Cin, k, k, k, seed=91, scale=(Cin * k ** 3) ** -0.5).bfloat16().float()
    b = rnd(Cout, seed=92, scale=0.1)
    A = (0.5 + torch.rand(B, Cin, generator=torch.Generator().manual_seed(93))).to(DEV)
    S = rnd(B, Cin, seed=94, scale=0.5)
    slope = torch.full((1,), 0.2, device=DEV)
    u = x * A[:, :, None, None, None] + S[:, :, None, None, None]
    xin = {"none": u, "relu": torch.relu(u), "leaky": torch.where(u > 0, u, 0.2 * u)}[act_name]
    ref = F.conv3d(xin.bfloat16().float(), w, b, padding=(k - 1) // 2)      # the materialised tensor is bf16 too
    xv = to_vol(x, dtype)
    if k == 1 and Cout == 1:        # per-sample weights, like reduce_channels
        wp = w.reshape(1, 1, Cout, Cin).expand(B, 1, Cout, Cin).contiguous().to(dtype)
        kw = dict(w_bstride=Cout * Cin, bias_bstride=0, impl=L.IMPL_SIMT)
        bias = b
    else:
        wp = ops.pack_weight(w, False, Cin, Cout, dtype)
        kw, bias = {}, b
    pro = ops.Deferred(xv, A, S, act, slope if act == L.ACT_LEAKY else None)
    probe_y = torch.empty(B, D, H, W, Cout, device=DEV, dtype=dtype)
    a = ops._conv_args(xv, wp, None, probe_y, ksize=k, stride=1, transposed=False, cout_comp=Cout,
                       w_bstride=kw.get("w_bstride", 0), impl=kw.get("impl", L.IMPL_AUTO))
    a.in_scale, a.in_shift, a.in_act = L.ptr(A), L.ptr(S), act
    a.in_slope = L.ptr(slope) if act == L.ACT_LEAKY else None
    assert bool(L.lib().coma_conv3d_prologue_supported(ctypes.byref(a))) == fused
    y, st = ops.conv_raw(pro, wp, bias, ksize=k, want_stats=True, **kw)
    assert err(to_ncdhw(y), ref) < 1.5e-2
    # against the same conv on the materialised input (same kernel family, so only the bf16 rounding point differs)
    ym, stm = ops.conv_raw(pro.materialize(), wp, bias, ksize=k, want_stats=True, **kw)
    assert err(y.float(), ym.float()) < 1.5e-2
    assert err(st.sum(dim=1), stm.sum(dim=1)) < 5e-3
    # the exact CUDA-core reference implementation of the prologue
    if not (k == 1 and Cout == 1):
        ys, _ = ops.conv_raw(pro, wp, bias, ksize=k, impl=L.IMPL_SIMT)
        assert err(y.float(), ys.float()) < 1.5e-2


# ---- BASELINE-size (batch 8 x 128^3) checks through size-independent properties ------------------------------------------
FULL = (8, 128, 128, 128)


def _grid_volume(B, D, H, W, Cn, seed):
    """bf16 values on a coarse grid (multiples of 1/4 in [-2, 2]): sums of two such volumes are exact in bf16."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randint(-8, 9, (B, D, H, W, Cn), device=DEV, generator=g).float() / 4).bfloat16()


@pytest.mark.parametrize("cin,cout", [(16, 16), (64, 32)])
def test_full_size_conv_is_linear_and_translation_equivariant(cin, cout):
    """Plane-ring tcgen05 conv at the benchmark size: conv(x1 + x2) = conv(x1) + conv(x2) up to output rounding, the
    fused statistics are linear and match the stored output, and shifting the input by one tile shifts the output
    bit-exactly (tiling, halo and segment bookkeeping do not depend on position)."""
    B, D, H, W = FULL
    w = (rnd(cout, cin, 3, 3, 3, seed=70, scale=(cin * 27) ** -0.5)).bfloat16().float()
    wp = ops.pack_weight(w, False, cin, cout, torch.bfloat16)
    x1, x2 = _grid_volume(B, D, H, W, cin, 71), _grid_volume(B, D, H, W, cin, 72)
    y1, s1 = ops.conv_raw(x1, wp, None, ksize=3, want_stats=True)
    y2, s2 = ops.conv_raw(x2, wp, None, ksize=3, want_stats=True)
    y12, s12 = ops.conv_raw(x1 + x2, wp, None, ksize=3, want_stats=True)
    scale = float(y12.float().abs().max())
    assert float((y12.float() - (y1.float() + y2.float())).abs().max()) < 1.5 * 2 ** -8 * scale      # three bf16 roundings
    lin = (s12.sum(dim=1)[..., 0] - s1.sum(dim=1)[..., 0] - s2.sum(dim=1)[..., 0]).abs().max()
    assert float(lin) < 1e-4 * float(s12.sum(dim=1)[..., 1].max()) ** 0.5 * (D * H * W) ** 0.5
    # checksum: per-(sample, channel) sums of the fp32 accumulators against the bf16 tensor that was stored
    stored = y12.float().sum(dim=(1, 2, 3))
    assert err(s12.sum(dim=1)[..., 0], stored) < 2e-3
    del y2, s1, s2, s12, x2
    # translation by one tile along w and h, one plane along d
    xs = torch.roll(x1, shifts=(1, 16, 8), dims=(1, 2, 3))
    ys, _ = ops.conv_raw(xs, wp, None, ksize=3)
    ref = torch.roll(y1, shifts=(1, 16, 8), dims=(1, 2, 3))
    inner = (slice(None), slice(2, D - 2), slice(18, H - 18), slice(10, W - 10))
    assert torch.equal(ys[inner], ref[inner])


def test_full_size_stride2_and_transposed_round_trip_shapes_and_adjointness():
    """<conv_s2(x), g> == <x, convT(g)> for the stride-2 pair (dgrad of one is the other) at the benchmark size: the
    tcgen05 stride-2 and transposed kernels are adjoint to bf16 accuracy."""
    B, D, H, W = FULL
    cin, cout = 32, 64
    w = rnd(cout, cin, 3, 3, 3, seed=73, scale=(cin * 27) ** -0.5).bfloat16().float()
    x = _grid_volume(B, D, H, W, cin, 74)
    g = _grid_volume(B, D // 2, H // 2, W // 2, cout, 75)
    y, _ = ops.conv_raw(x, ops.pack_weight(w, False, cin, cout, torch.bfloat16), None, ksize=3, stride=2)
    # adjoint of conv(stride 2) = convT(stride 2) with the same kernel read as ConvTranspose3d weights [Cin=cout, Cout=cin]
    xt, _ = ops.conv_raw(g, ops.pack_weight(w, True, cout, cin, torch.bfloat16), None, ksize=3, stride=2, transposed=True)
    lhs = float((y.float() * g.float()).sum())
    rhs = float((x.float() * xt.float()).sum())
    assert abs(lhs - rhs) < 2e-3 * max(abs(lhs), abs(rhs), 1.0) + 2e-3 * float(y.float().norm() * g.float().norm()) * 1e-2


FULL_NUMERIC = [
    # (B, Cin, Cout, D, k, stride, transposed): the layers of the benchmarked network at batch 8
    (8, 64, 32, 128, 3, 1, False),     # merge conv of the top level (CTA-pair kernel)
    (8, 32, 32, 128, 3, 1, False),     # head conv 2 (plane-ring kernel)
    (8, 16, 16, 128, 3, 1, False),     # modulator stack conv
    (8, 32, 64, 128, 3, 2, False),     # first down-sampling conv (stride-2 plane ring)
    (8, 64, 32, 64, 3, 2, True),       # top-level transposed conv
    (8, 128, 64, 64, 3, 1, False),     # second-level merge conv (CTA pair, two K chunks)
]


@pytest.mark.parametrize("case", FULL_NUMERIC)
def test_full_size_tcgen05_conv_matches_cudnn_fp32(case):
    """The tcgen05 kernels at the BENCHMARK size, numerically: forward, data gradient and weight gradient against
    F.conv3d / F.conv_transpose3d in fp32 (TF32 off) on the same bf16-representable inputs.  Only accumulation order and
    the bf16 rounding of the stored result differ, so the bound is bf16's 2^-8 of full scale for tensors stored in bf16
    and 1e-3 for the fp32 weight gradient (1.7e7 voxels x 8 samples of fp32 accumulation, atomically combined)."""
    B, Cin, Cout, D, k, s, tr = case
    Di = D
    g = torch.Generator(device=DEV).manual_seed(90)
    x = torch.randn(B, Cin, Di, Di, Di, device=DEV, generator=g).bfloat16().float()
    wshape = (Cin, Cout, k, k, k) if tr else (Cout, Cin, k, k, k)
    w = (torch.randn(*wshape, device=DEV, generator=g) * (Cin * k ** 3) ** -0.5).bfloat16().float()
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    pad = (k - 1) // 2
    ref = F.conv_transpose3d(xr, wr, None, stride=s, padding=pad, output_padding=s - 1) if tr else F.conv3d(xr, wr, None, stride=s, padding=pad)
    gy = torch.randn(ref.shape, device=DEV, generator=g).bfloat16().float()
    ref.backward(gy)
    ref = ref.detach()
    xv = x.permute(0, 2, 3, 4, 1).contiguous().bfloat16().requires_grad_(True)
    del x
    wp = w.clone().requires_grad_(True)
    y, _ = ops.conv3d(xv, wp, None, ops.ConvCfg(ksize=k, stride=s, transposed=tr))
    assert err(y.permute(0, 4, 1, 2, 3), ref) < 2 ** -8
    del ref
    y.backward(gy.permute(0, 2, 3, 4, 1).contiguous().bfloat16())
    assert err(xv.grad.permute(0, 4, 1, 2, 3), xr.grad) < 2 ** -8
    assert err(wp.grad, wr.grad) < 1e-3


@pytest.mark.parametrize("case", [(8, 128, 64, 64, 3, 1, False), (8, 64, 32, 128, 3, 1, False), (8, 16, 16, 128, 3, 1, False),
                                  (8, 32, 64, 128, 3, 2, False), (8, 64, 32, 64, 3, 2, True)])
def test_warp_specialised_convs_are_run_to_run_identical(case):
    """A missed barrier or phase slip in the producer / MMA / epilogue pipelines (CTA pair, plane ring, stride 2, transposed)
    shows up as run-to-run differences: outputs and fused statistics must be bit-identical over repeated launches at the
    benchmark size."""
    B, Cin, Cout, D, k, s, tr = case
    x = rnd(B, D, D, D, Cin, seed=80).bfloat16()
    wshape = (Cin, Cout, k, k, k) if tr else (Cout, Cin, k, k, k)
    wp = ops.pack_weight(rnd(*wshape, seed=81, scale=(Cin * 27) ** -0.5), tr, Cin, Cout, torch.bfloat16)
    y0, s0 = ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=tr, want_stats=True)
    y0, s0 = y0.clone(), s0.sum(dim=1).clone()
    for _ in range(12):
        y, st = ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=tr, want_stats=True)
        assert torch.equal(y, y0) and torch.equal(st.sum(dim=1), s0)


@pytest.mark.parametrize("case", [(4, 32, 64, 128, 1), (4, 16, 16, 128, 1), (4, 64, 32, 64, 2), (4, 256, 128, 16, 2), (4, 1, 16, 128, 1)])
def test_tcgen05_weight_gradients_are_bit_identical_from_run_to_run(case):
    """the tcgen05 weight gradients no longer drain through fp32 atomics -- every CTA writes its partial
    [27][Cg][Cx] block to the caller's workspace and a second kernel sums the blocks in CTA order (coma_wgrad_args.workspace), so
    repeated launches at the benchmark size give bit-identical gradients (stride 1 and the stride-2 kernel)."""
    B, Cg, Cx, Dg, s = case
    g = rnd(B, Dg, Dg, Dg, Cg, seed=83).bfloat16()
    x = rnd(B, Dg * s, Dg * s, Dg * s, Cx, seed=84).bfloat16()
    assert ops.DETERMINISTIC_WGRAD
    first = ops.wgrad_raw(g, x, ksize=3, stride=s).clone()
    assert float(first.abs().max()) > 0
    for _ in range(8):
        assert torch.equal(ops.wgrad_raw(g, x, ksize=3, stride=s), first)


# ---- input preparation: coma_prepare_volumes against oracle/prepare.py, bit for bit ----
@pytest.mark.parametrize("shape,spacing,pad_dims,resize", [
    ((24, 30, 20), (1.0, 1.0, 1.0), (16, 16, 16), True),      # 1 mm -> 2 mm, centred padding, odd pads
    ((24, 40, 20), (1.0, 1.0, 1.0), (16, 16, 16), True),      # y longer than the target: cropped at its end
    ((17, 19, 23), (1.2, 0.9, 1.5), (20, 20, 20), True),      # anisotropic, half-integer source coordinates
    ((9, 7, 11), (4.0, 4.0, 4.0), (24, 24, 24), True),        # coarse -> fine: the last slab falls outside (default value)
    ((12, 10, 14), (1.0, 1.0, 1.0), (16, 16, 16), False),     # no resample, padding only
    ((16, 12, 14), (2.0, 2.0, 2.0), (16, 16, 16), True),      # z already at the target: apply_transforms skips the padding
])
def test_prepare_volumes_matches_oracle(shape, spacing, pad_dims, resize):
    import numpy as np
    from coma_unet_b200 import prepare_volumes
    from oracle import prepare as oprep
    g = torch.Generator().manual_seed(sum(shape))
    mri = torch.rand(shape, generator=g)
    tau = torch.randn(shape, generator=g)
    roi = torch.randint(0, 4, shape, generator=g).float() * 1007.0
    mri.view(-1)[::37] = float("nan"); tau.view(-1)[::41] = float("inf"); tau.view(-1)[::43] = float("-inf")
    want = oprep.prepare_sample(mri.numpy(), tau.numpy(), roi.numpy(), spacing, resize, pad_dims)
    got = prepare_volumes(mri.to(DEV), tau.to(DEV), roi.to(DEV), spacing, resize, pad_dims)
    for w, t, name in zip(want, got, ("mri", "tau", "roi")):
        assert tuple(t.shape) == tuple(w.shape), name
        assert np.array_equal(t.cpu().numpy().view(np.uint32), w.numpy().view(np.uint32)), name
    only_roi = prepare_volumes(None, None, roi.to(DEV), spacing, resize, pad_dims)
    assert only_roi[0] is None and only_roi[1] is None and torch.equal(only_roi[2], got[2])


def test_prepare_volumes_full_size_properties():
    """BASELINE-size case (a 1 mm 256^3 scan -> the 128^3 the model takes): every other voxel survives, the MRI is zero wherever
    the ROI map is, the call is idempotent on its own output, and CPU tensors are refused."""
    from coma_unet_b200 import prepare_volumes
    g = torch.Generator(device=DEV).manual_seed(5)
    mri = torch.rand((256, 256, 256), generator=g, device=DEV)
    roi = (torch.rand((256, 256, 256), generator=g, device=DEV) > 0.3).float() * 17.0
    m, _, r = prepare_volumes(mri, None, roi, (1.0, 1.0, 1.0))
    assert m.shape == (1, 128, 128, 128) and torch.equal(r[0], roi[::2, ::2, ::2])
    assert torch.equal(m[0], mri[::2, ::2, ::2] * (roi[::2, ::2, ::2] != 0))
    m2, _, r2 = prepare_volumes(m[0], None, r[0], (2.0, 2.0, 2.0))
    assert torch.equal(m2, m) and torch.equal(r2, r)
    with pytest.raises(RuntimeError):
        prepare_volumes(mri.cpu(), None, None)


# ------------------------------------------------------------------------------------------------
# coma_weight_layout: one launch instead of the permute / pad / cast / flip chain (bit-exact: it only moves and rounds)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cout,cin,k,transposed,stride", [
    (16, 1, 3, False, 1), (32, 16, 3, False, 1), (64, 32, 3, False, 2), (32, 64, 3, True, 2), (1, 16, 1, False, 1),
    (16, 32, 1, False, 1), (256, 128, 3, False, 1), (20, 7, 3, False, 1), (7, 20, 3, True, 2), (1, 16, 3, False, 1), (24, 40, 2, False, 1)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_weight_layout_kernel_matches_the_tensor_op_chain(monkeypatch, cout, cin, k, transposed, stride, dtype):
    shape = (cin, cout, k, k, k) if transposed else (cout, cin, k, k, k)
    w = torch.nn.Parameter(rnd(*shape, seed=cout * 131 + cin))
    cin_buf, cout_comp = -(-cin // 16) * 16, -(-cout // 16) * 16
    cout_store = cout_comp if cout > 4 else cout          # narrow outputs are stored unpadded (the dy of the data gradient)
    got_fwd = ops.pack_weight(w, transposed, cin_buf, cout_comp, dtype)
    got_adj = ops.pack_adjoint(w, transposed, stride, cin_buf, cout_comp, cout_store, dtype)
    rows, cols = (cin_buf, cout_store) if transposed else (cout_store, cin_buf)
    dwp = rnd(k ** 3, rows, cols, seed=5)
    got_dw = ops.unpack_weight_gradient(dwp, w)
    monkeypatch.setattr(ops, "PACK_KERNEL", False)
    w.__dict__.pop("_coma_packed", None)
    want_fwd = ops.pack_weight(w, transposed, cin_buf, cout_comp, dtype)
    want_adj = ops.pack_adjoint(w, transposed, stride, cin_buf, cout_comp, cout_store, dtype)
    want_dw = ops.unpack_weight_gradient(dwp, w)
    for got, want in ((got_fwd, want_fwd), (got_adj, want_adj), (got_dw, want_dw)):
        assert got.shape == want.shape and got.dtype == want.dtype and got.is_contiguous()
        assert torch.equal(got, want)


# ------------------------------------------------------------------------------------------------
# coma_film_mlp_fwd / _bwd: every FiLM MLP of a model in one launch each way, against the per-layer nn.Sequential chain
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [1, 4, 7])
def test_fused_film_mlps_match_the_per_layer_modules(B):
    torch.manual_seed(11)
    layers = [(5, 16), (6, 32), (6, 256), (5, 1), (6, 24)]          # (num_covars, out_channels)
    mods = [torch.nn.Sequential(torch.nn.Linear(n, 64), torch.nn.ReLU(), torch.nn.Linear(64, 2 * c)).to(DEV) for n, c in layers]
    cov = rnd(B, 6, seed=3)
    params = [p for m in mods for p in (m[0].weight, m[0].bias, m[2].weight, m[2].bias)]
    outs = ops.FilmAllFn.apply(cov, tuple(layers), *params)
    assert len(outs) == 2 * len(layers)
    gs = [rnd(B, c, seed=20 + l) for l, (_, c) in enumerate(layers)]
    hs = [rnd(B, c, seed=40 + l) for l, (_, c) in enumerate(layers)]
    skip = 3                                                         # a layer whose outputs nobody consumes keeps grad None
    loss = sum((outs[2 * l] * gs[l]).sum() + (outs[2 * l + 1] * hs[l]).sum() for l in range(len(layers)) if l != skip)
    loss.backward()
    got = [None if p.grad is None else p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    for l, (m, (n, c)) in enumerate(zip(mods, layers)):
        dgamma, beta = m(cov[:, :n]).chunk(2, dim=-1)
        assert outs[2 * l].is_contiguous() and outs[2 * l].shape == (B, c)
        # fp32 sums in a different order; absolute floor because a one-element tensor (B = 1, C = 1) can sit next to zero
        close = lambda a, b: torch.allclose(a, b, rtol=1e-4, atol=1e-5)      # noqa: E731
        assert close(outs[2 * l], dgamma) and close(outs[2 * l + 1], beta), (l, (outs[2 * l] - dgamma).abs().max())
        if l != skip:
            ((dgamma * gs[l]).sum() + (beta * hs[l]).sum()).backward()
    for i, p in enumerate(params):
        if i // 4 == skip:
            assert got[i] is None
        else:
            assert got[i].shape == p.grad.shape
            assert torch.allclose(got[i], p.grad, rtol=1e-4, atol=1e-5), (i, (got[i] - p.grad).abs().max(), p.grad.abs().max())
