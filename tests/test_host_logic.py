"""CPU: host-side logic of the product package (no kernels run)."""
import pytest
import torch

import coma_unet_b200 as cu
from coma_unet_b200 import ops
from oracle import criterions as ocrit
from oracle import model as omodel
from tests.golden import common


def test_state_dict_keys_match_oracle_and_reference_layout():
    kw = dict(latent_spaces=[2048] * 5, conditional=True, prompt_shape=(16, 16, 16))
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, [8, 16, 32, 64, 128], [2] * 5, **kw)
    o = omodel.ContrastiveAttentionUNET_DP(3, 1, 1, [8, 16, 32, 64, 128], [2] * 5, **kw)
    assert set(m.state_dict()) == set(o.state_dict())
    m.load_state_dict(o.state_dict(), strict=True)
    for key in ["model.0.conv.0.conv.weight", "model.1.submodule.1.attention.W_g.0.conv.weight",
                "model.1.upconv.up.conv.weight", "model.1.merge.adn.A.weight", "projection_heads.0.conv.conv.1.adn.N.bias",
                "pos_dynamic_prompt", "model.2.routing.weight", "model.0.conv.1.film.2.weight"]:
        assert key in m.state_dict(), key
    # non-conditional variant builds too
    p = cu.ContrastiveAttentionUNET_DP(3, 1, 1, [8, 16, 32, 64, 128], [2] * 5, latent_spaces=[2048] * 5,
                                       conditional=False, prompt_shape=(16, 16, 16))
    q = omodel.ContrastiveAttentionUNET_DP(3, 1, 1, [8, 16, 32, 64, 128], [2] * 5, latent_spaces=[2048] * 5,
                                           conditional=False, prompt_shape=(16, 16, 16))
    assert set(p.state_dict()) == set(q.state_dict())


def test_public_surface():
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, [8, 16, 32, 64, 128], [2] * 5, latent_spaces=[2048] * 5,
                                       conditional=True, prompt_shape=(16, 16, 16))
    assert m.get_depth() == 5 and len(m.roi_indices) == 36 and len(m.roi_names) == 36
    assert m.roi_ind_names_dict[1001] == "ctx-lh-bankssts" and m.roi_ind_vol_names_dict[17] == "vol_Left_Hippocampus"
    assert m.all_stages and not m.only_stage_two and not m.with_uq and not m.decoder_ds and not m.embeddings_out
    m.set_training(False)
    assert m.training is False
    m.set_save_attn(None)
    assert m.model[1].save_attn is None and m.model[1].attention.save_attn is None


def test_pack_weight_layout():
    w = torch.arange(2 * 3 * 27, dtype=torch.float32).reshape(2, 3, 3, 3, 3)   # Conv3d [Cout=2, Cin=3, k,k,k]
    p = ops.pack_weight(w, False, 3, 2, torch.float32)
    assert p.shape == (27, 2, 3)
    assert p[5, 1, 2] == w[1, 2, 0, 1, 2]            # tap 5 = (kd=0, kh=1, kw=2)
    p16 = ops.pack_weight(w, False, 16, 16, torch.bfloat16)
    assert p16.shape == (27, 16, 16) and p16[:, 2:, :].abs().sum() == 0 and p16[:, :, 3:].abs().sum() == 0
    wt = torch.arange(3 * 2 * 27, dtype=torch.float32).reshape(3, 2, 3, 3, 3)  # ConvTranspose3d [Cin=3, Cout=2, ...]
    pt = ops.pack_weight(wt, True, 3, 2, torch.float32)
    assert pt.shape == (27, 2, 3) and pt[13, 1, 2] == wt[2, 1, 1, 1, 1]


def test_vol_strides():
    buf = torch.zeros(2, 4, 4, 4, 32)
    assert ops.vol_cs(buf) == 32 and ops.vol_cs(buf[..., 16:]) == 32 and ops.vol_cs(buf[..., :8]) == 32
    x = torch.zeros(2, 1, 4, 4, 4)
    assert ops.vol_cs(ops.ncdhw_to_vol(x, torch.float32)) == 1
    with pytest.raises(AssertionError):
        ops.vol_cs(torch.zeros(2, 32, 4, 4, 4).permute(0, 2, 3, 1, 4))


def test_rnc_loss_matches_oracle_on_cpu():
    g = torch.Generator().manual_seed(0)
    f = torch.randn(5, 32, generator=g, requires_grad=True)
    y = torch.rand(5, 6, generator=g)
    a = cu.RnCLoss()(f, y)
    f2 = f.detach().clone().requires_grad_(True)
    b = ocrit.RnCLoss()(f2, y)
    a.backward()
    b.backward()
    assert torch.allclose(a, b) and torch.allclose(f.grad, f2.grad)
    assert cu.RnCLoss()(f[:1], y[:1]) == 0.0


def test_product_has_no_cpu_fallback():
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, [8, 16, 32, 64, 128], [2] * 5, latent_spaces=[2048] * 5,
                                       conditional=True, prompt_shape=(16, 16, 16)).eval()
    mri, tau, roi, covars, dicts = common.synthetic_batch(1, (16, 16, 16), 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    with pytest.raises(RuntimeError, match="CUDA"):
        cu.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)(mri, tau, roi)


def test_synthetic_dataset_contract():
    ds = cu.SyntheticVolumeDataset(length=3, shape=(16, 16, 16))
    mri, tau, roi, (abeta, covars), path = ds[1]
    assert mri.shape == tau.shape == roi.shape == (1, 16, 16, 16) and mri.dtype == torch.float32
    assert covars.shape == (1, 6) and covars.dtype == torch.float64 and isinstance(path, str)
    assert cu.SyntheticVolumeDataset(length=1, shape=(8, 8, 8), flavour="adni")[0][3][1].dtype == torch.float32
    from torch.utils.data import DataLoader
    batch = next(iter(DataLoader(ds, batch_size=2)))
    assert batch[0].shape == (2, 1, 16, 16, 16) and batch[3][1].shape == (2, 1, 6) and len(batch[4]) == 2
    assert set(ds.roi_predictions(0)) == set(common.roi_names())


def test_prepare_geometry_follows_the_reference_padding_rules():
    """Host integers of the GPU input preparation: numpy round-half-even output size (VolumeDataset.py:241-245), centred padding with
    the odd voxel at the end, only the y axis cropped (data_util.py:814-828), no padding when the z extent already matches
    (VolumeDataset.py:261-264)."""
    from coma_unet_b200 import prepare_geometry
    res, out, before, ratio = prepare_geometry((255, 256, 250), (1.0, 1.0, 1.0))
    assert res == [128, 128, 125] and out == res and before == [0, 0, 0] and ratio == [2.0, 2.0, 2.0]   # round(127.5) = 128; z matches: x stays 125
    res, out, before, _ = prepare_geometry((250, 256, 251), (1.0, 1.0, 1.0))
    assert res == [125, 128, 126] and out == [128, 128, 128] and before == [1, 0, 1]                      # round(125.5) = 126; odd voxel at the end
    res, out, before, _ = prepare_geometry((100, 300, 90), (1.0, 1.0, 1.0))
    assert res == [50, 150, 45] and out == [128, 128, 128] and before == [39, 0, 41]                               # y cropped at its end
    res, out, before, _ = prepare_geometry((256, 200, 200), (1.0, 1.0, 1.0))
    assert res == [128, 100, 100] and out == res and before == [0, 0, 0]                                           # z matches: no transform
    res, out, _, ratio = prepare_geometry((64, 64, 64), (1.0, 1.0, 1.0), resize=False, pad_dims=None)
    assert res == [64, 64, 64] and out == res and ratio == [1.0, 1.0, 1.0]


def test_packed_weight_cache_follows_the_weight_epoch():
    """Fused optimizers update parameters in place without bumping ``_version`` (round-2 finding): the packed-weight cache is
    keyed by the weight epoch as well, which every training-mode forward and every train()/eval() switch bumps."""
    w = torch.nn.Parameter(torch.arange(2 * 3 * 27, dtype=torch.float32).reshape(2, 3, 3, 3, 3))
    p0 = ops.pack_weight(w, False, 3, 2, torch.float32)
    assert ops.pack_weight(w, False, 3, 2, torch.float32) is p0                 # cached
    w.data.mul_(2.0)                                                               # what a fused optimizer step looks like: no version bump
    assert ops.pack_weight(w, False, 3, 2, torch.float32) is p0                 # version-only key: stale by construction ...
    ops.bump_weight_epoch()
    p1 = ops.pack_weight(w, False, 3, 2, torch.float32)                         # ... the epoch refreshes it
    assert p1 is not p0 and torch.equal(p1, 2 * p0)
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, [8, 16, 32, 64, 128], [2] * 5, latent_spaces=[2048] * 5, conditional=True,
                                       prompt_shape=(16, 16, 16))
    e0 = ops.weight_epoch()
    m.train(True)
    m.eval()
    m.set_training(False)
    assert ops.weight_epoch() >= e0 + 3
    ops.invalidate_weight_caches(m)
    assert ops.weight_epoch() > e0 + 3


def test_roi_table_and_prompt_declarations():
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, [8, 16, 32, 64, 128], [2] * 5, latent_spaces=[2048] * 5, conditional=True,
                                       prompt_shape=(16, 16, 16))
    _, _, _, _, dicts = common.synthetic_batch(2, (16, 16, 16), 3)
    dicts[1][m.roi_names[4]]["loc"] = float("nan")
    lut = m.roi_lut_host(dicts)
    assert lut.shape == (2, 36, 2) and lut.dtype.name == "float32"
    assert lut[0, 0, 0] == pytest.approx(dicts[0][m.roi_names[0]]["loc"]) and lut[1, 4, 0] == 0.0      # nan_to_num, :641
    dd = m.data_dependent_parameters()
    assert len(dd) == 2 and dd[0] is m.pos_dynamic_prompt and dd[1] is m.neg_dynamic_prompt
    # the FiLM MLPs (evaluated by one fused launch each way under autograd) are declared to the data-parallel engine as
    # "gradient arrives at the end of backward": every film parameter, nothing else, none of the data-dependent ones
    late = m.end_of_backward_parameters()
    film = [p for name, p in m.named_parameters() if ".film." in name]
    assert film and {id(p) for p in late} == {id(p) for p in film}
    assert not {id(p) for p in late} & {id(p) for p in dd}


def _layout_as_documented(param, R_pad, C_pad, swap, flip):
    """include/coma_b200.h, coma_weight_layout: packed[t][r][c] = param[a][b][t'] with (a, b) = (r, c) or, with swap, (c, r),
    t' = T - 1 - t with flip, zero beyond A / B."""
    import numpy as np
    A, B, T = param.shape
    out = np.zeros((T, R_pad, C_pad), param.dtype)
    for t in range(T):
        ts = T - 1 - t if flip else t
        blk = param[:, :, ts].T if swap else param[:, :, ts]
        out[t, :min(R_pad, blk.shape[0]), :min(C_pad, blk.shape[1])] = blk[:R_pad, :C_pad]
    return out


@pytest.mark.parametrize("transposed,stride", [(False, 1), (False, 2), (True, 2)])
def test_weight_layout_abi_documentation_matches_the_tensor_op_chain(transposed, stride):
    """The swap / flip table in the header against ops.pack_weight / pack_adjoint / unpack_weight_gradient on CPU tensors (the
    chain of tensor ops; the GPU suite pins the kernel to the same chain bit for bit)."""
    import numpy as np
    torch.manual_seed(0)
    cout, cin, k = 5, 3, 3
    w = torch.randn((cin, cout, k, k, k) if transposed else (cout, cin, k, k, k))
    flat = w.reshape(w.shape[0], w.shape[1], -1).numpy()
    fwd = ops.pack_weight(w, transposed, 16, 16, torch.float32).numpy()
    assert np.array_equal(fwd, _layout_as_documented(flat, 16, 16, swap=transposed, flip=False))
    adj = ops.pack_adjoint(w, transposed, stride, 16, 16, 16, torch.float32).numpy()
    assert np.array_equal(adj, _layout_as_documented(flat, 16, 16, swap=not transposed, flip=(not transposed and stride == 1)))
    dwp = torch.randn(k ** 3, 16, 16)
    dw = ops.unpack_weight_gradient(dwp, w)
    assert dw.shape == w.shape
    A, B = w.shape[0], w.shape[1]
    assert np.array_equal(dw.reshape(A, B, -1).numpy(), dwp[:, :A, :B].permute(1, 2, 0).numpy())
