"""Oracle restatement of the reference model classes.  Test infrastructure.

Follows attn_unet_data_parallel.py:120-693 (plain PyTorch, fp32, NCDHW, CPU-runnable):

* ``UpBlock``                       :120-131
* ``ObservableAttentionBlock``      :134-150
* ``AttentionLayer``                :152-240   (nested-tuple plumbing restated as two lists)
* ``ObservableAttentionUnet``       :243-432
* ``ProjectionHead``                :436-454
* ``StackedFusionConvLayers``       :480-501
* ``ContrastiveAttentionUNET_DP``   :503-693

Pinned by tests/golden/: the reference's own classes are imported from /root/reference
(with ``oracle.monai_blocks`` / ``oracle.cond_conv`` injected for the missing ``monai`` /
``CondConv`` modules) and their outputs on seeded inputs are the fixtures this file is
checked against (tests/test_oracle_golden.py).

One extension, passed through the constructor's ``**kwargs`` (:520,620): ``prompt_shape``
(default ``(128,128,128)``, the reference's hard-coded prompt size, :544-545,610).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.nn as nn

from . import cond_conv
from . import monai_blocks as mb

ROI_INDICES = [
    1001, 1006, 1007, 1009, 1015, 1016, 1030, 1034, 1033, 1008, 1025, 1029, 1031, 1022, 17, 18,
    2001, 2006, 2007, 2009, 2015, 2016, 2030, 2034, 2033, 2008, 2025, 2029, 2031, 2022, 49, 50, 51, 52, 53, 54,
]  # :561-564

_CTX = ["bankssts", "entorhinal", "fusiform", "inferiortemporal", "middletemporal", "parahippocampal",
        "superiortemporal", "transversetemporal", "temporalpole", "inferiorparietal", "precuneus",
        "superiorparietal", "supramarginal", "postcentral"]
ROI_NAMES = ([f"ctx-lh-{n}" for n in _CTX] + ["Left-Hippocampus", "Left-Amygdala"]
             + [f"ctx-rh-{n}" for n in _CTX]
             + ["Right-Thalamus-Proper", "Right-Caudate", "Right-Putamen", "Right-Pallidum",
                "Right-Hippocampus", "Right-Amygdala"])  # :567-578, same order as ROI_INDICES


def save_attention_coeffs(path: str, coeff: torch.Tensor) -> None:
    """Stand-in for data_util.save_attention_coeffs (data_util.py:802-811; NIfTI there, .npy here)."""
    vol = np.squeeze(coeff.detach().cpu().numpy())
    stem = path.rsplit(".", 1)[0] if "." in path else path
    np.save(f"{stem}_vdim{vol.shape[-1]}.npy", vol)


class UpBlock(mb.UpConv):
    def __init__(self, conditional, num_covars=0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if conditional:
            self.up = cond_conv.CondConvolution(dropout=0.0, is_transposed=True, num_covars=num_covars, *args, **kwargs)
        self.conditional = conditional

    def forward(self, x, covariate=None):
        return self.up(x, covariate) if self.conditional else self.up(x)


class ObservableAttentionBlock(mb.AttentionBlock):
    save_attn = None

    def forward(self, g, x):
        psi = self.psi(self.relu(self.W_g(g) + self.W_x(x)))
        return (x * psi, psi) if self.save_attn else x * psi


class AttentionLayer(mb.AttentionLayer):
    save_attn = None

    def __init__(self, spatial_dims, in_channels, out_channels, submodule, up_kernel_size=3, strides=2,
                 dropout=0.0, conditional=False, num_covars=0):
        super().__init__(spatial_dims, in_channels, out_channels, submodule, up_kernel_size, strides, dropout)
        self.attention = ObservableAttentionBlock(spatial_dims, f_g=in_channels, f_l=in_channels,
                                                  f_int=in_channels // 2)
        self.upconv = UpBlock(conditional=conditional, spatial_dims=spatial_dims, in_channels=out_channels,
                              out_channels=in_channels, strides=strides, kernel_size=up_kernel_size,
                              num_covars=num_covars)

    def set_save_attn(self, status):
        self.save_attn = status
        self.attention.save_attn = status

    def forward(self, x, covariate=None):
        """Returns (decoder output at this level, encoder tensors from here down, decoder tensors from here down).

        The reference returns the nested tuple ``att_m, (x, (x_sub, rest))`` (:240) and unrolls it in
        ``ObservableAttentionUnet.forward`` (:399-421); the two lists are that unrolling.
        """
        cov5 = covariate[:, :, :5] if covariate is not None else None  # :209,212
        if isinstance(self.submodule, nn.Sequential):
            block, deeper = self.submodule[0], self.submodule[1]
            x_sub, encs, decs = deeper(block(x, covariate=cov5), covariate=covariate)
        else:
            x_sub = self.submodule(x, covariate=cov5)
            encs, decs = [x_sub], []
        fromlower = self.upconv(x_sub, covariate)
        att = self.attention(g=fromlower, x=x)
        if self.save_attn is not None:
            att, coeff = att
            save_attention_coeffs(self.save_attn, coeff)
        att_m = self.merge(torch.cat((att, fromlower), dim=1))
        return att_m, [x] + encs, [att_m] + decs


class _PlainBlock(mb.ConvBlock):
    def __init__(self, num_covars=0, **kw):
        super().__init__(**kw)

    def forward(self, x, covariate=None):
        return super().forward(x)


class _PlainConvolution(mb.Convolution):
    def __init__(self, *a, num_experts=1, num_covars=0, **kw):
        super().__init__(*a, **kw)

    def forward(self, x, covariate=None):
        return super().forward(x)


class ObservableAttentionUnet(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, channels: Sequence[int],
                 strides: Sequence[int], kernel_size=3, up_kernel_size=3, dropout: float = 0.0,
                 conditional: bool = False):
        super().__init__()
        self.dimensions, self.in_channels, self.out_channels = spatial_dims, in_channels, out_channels
        self.channels, self.strides, self.kernel_size = channels, strides, kernel_size
        self.dropout, self.conditional, self.up_kernel_size = dropout, conditional, up_kernel_size
        self.with_regression = True
        self.save_attn = None
        # The non-conditional reference path would call MONAI blocks with a ``covariate`` kwarg and
        # fail (:428); the plain wrappers accept and ignore it so conditional=False is usable.
        Block = cond_conv.CondConvBlock if conditional else _PlainBlock
        Conv = cond_conv.CondConvolution if conditional else _PlainConvolution
        ncov_up = 5 + int(self.with_regression)

        head = Block(spatial_dims=spatial_dims, in_channels=in_channels, out_channels=channels[0],
                     dropout=dropout, num_covars=5)
        reduce_channels = Conv(spatial_dims=spatial_dims, in_channels=channels[0], out_channels=out_channels,
                               kernel_size=1, strides=1, padding=0, conv_only=True, num_experts=8,
                               num_covars=ncov_up)

        def level(ch, st):
            down = Block(spatial_dims=spatial_dims, in_channels=ch[0], out_channels=ch[1], strides=st[0],
                         dropout=dropout, num_covars=5)
            sub = nn.Sequential(down, level(ch[1:], st[1:])) if len(ch) > 2 else down
            return AttentionLayer(spatial_dims=spatial_dims, in_channels=ch[0], out_channels=ch[1], submodule=sub,
                                  up_kernel_size=up_kernel_size, strides=st[0], dropout=dropout,
                                  conditional=conditional, num_covars=ncov_up)

        self.model = nn.ModuleList([head, level(list(channels), list(strides)), reduce_channels])

    def set_save_attn(self, value):
        layer = self.model[1]
        while isinstance(layer, AttentionLayer):
            layer.set_save_attn(value)
            sub = layer.submodule
            if not isinstance(sub, nn.Sequential):
                break
            layer = sub[-1]

    def forward(self, x, covariate=None):
        head, encdec, reduce_channels = self.model
        x = head(x, covariate=covariate[:, :, :5] if covariate is not None else None)  # :428
        x, encoder_extractions, decoder_extractions = encdec(x, covariate)              # :397-421
        x = reduce_channels(x, covariate=covariate)                                     # :425
        return x, encoder_extractions, decoder_extractions


class ProjectionHead(nn.Module):
    def __init__(self, in_channels, out_channels, latent_space_dim, kernel_size=3):
        super().__init__()
        self.conv = mb.ConvBlock(3, in_channels, 1, kernel_size=1)
        self.act_fn = nn.ReLU()

    def forward(self, x):
        return self.act_fn(self.conv(x).flatten(1))


class StackedFusionConvLayers(nn.Module):
    def __init__(self, input_feature_channels, bottleneck_feature_channel, output_feature_channels, num_convs,
                 nonlin=nn.LeakyReLU, nonlin_kwargs=None):
        super().__init__()
        self.input_channels, self.output_channels = input_feature_channels, output_feature_channels
        act = (nonlin, nonlin_kwargs or {"negative_slope": 1e-2, "inplace": True})
        widths = [input_feature_channels] + [bottleneck_feature_channel] * (num_convs - 1) + [output_feature_channels]
        self.blocks = nn.Sequential(*[mb.Convolution(3, widths[i], widths[i + 1], act=act) for i in range(num_convs)])

    def forward(self, x):
        return self.blocks(x)


class ContrastiveAttentionUNET_DP(ObservableAttentionUnet):
    def __init__(self, spatial_dims, in_channels, out_channels, channels, strides, latent_spaces, kernel_size=3,
                 up_kernel_size=3, dropout=0, training=True, embeddings_out=False, conditional=False,
                 decoder_ds=False, **kwargs):
        super().__init__(spatial_dims, in_channels, out_channels, channels, strides, kernel_size, up_kernel_size,
                         dropout, conditional)
        self.training = training
        self.embeddings_out, self.decoder_ds = embeddings_out, decoder_ds
        self.depth = len(channels)
        ps = tuple(kwargs.get("prompt_shape", (128, 128, 128)))

        self.projection_heads = nn.ModuleList([
            ProjectionHead(channels[i], int((128 / 2 ** i) ** 3), latent_spaces[i]) for i in range(len(channels))])
        self.final_projection_head = nn.Sequential(nn.AdaptiveAvgPool3d(1), nn.Linear(out_channels, latent_spaces[-1]),
                                                   nn.ReLU())
        self.pos_dynamic_prompt = nn.Parameter(torch.randn(1, 1, *ps))
        self.neg_dynamic_prompt = nn.Parameter(torch.randn(1, 1, *ps))
        self.fusion_layer = StackedFusionConvLayers(2, 8, 1, num_convs=3)
        self.modulator = mb.Convolution(3, 2, 1, act="ReLU")       # never used in forward
        self.modulator_3c = mb.Convolution(3, 3, 1, act="ReLU")    # never used in forward
        self.reweigh = nn.Parameter(torch.ones(ps))                 # never used in forward
        self.final_act = nn.ReLU()
        self.pos_reweigh = nn.Parameter(torch.ones((1, *ps)))       # never used in forward
        self.neg_reweigh = nn.Parameter(torch.ones((1, *ps)))       # never used in forward
        self.deep_modulator_3c = StackedFusionConvLayers(3, 16, 1, num_convs=3)
        self.final_pred_head = mb.Convolution(3, 2, 1, kernel_size=1)

        self.roi_indices = list(ROI_INDICES)
        self.roi_names = list(ROI_NAMES)
        self.roi_ind_names_dict = dict(zip(ROI_INDICES, ROI_NAMES))
        self.roi_ind_vol_names_dict = {k: "vol_" + "_".join(v.split("-")) for k, v in self.roi_ind_names_dict.items()}
        self.general_dynamic_prompt = nn.Parameter(torch.randn(1, 1, *ps))
        self.roi_wise_reweigh = nn.ParameterList([nn.Parameter(torch.ones(1)) for _ in ROI_INDICES])  # unused
        self.all_stages, self.only_stage_two = True, False
        self.with_uq = kwargs.get("with_uq", False)

    def set_training(self, mode):
        self.training = mode

    def get_depth(self):
        return self.depth

    def forward_modulator_with_uq(self, x, out, covariate=None, roi_pred_dicts=None, sample_roi_mask=None):
        suvr = torch.zeros_like(out)
        saliency = torch.zeros_like(out)
        prompts = []
        for b in range(x.size(0)):
            positive = covariate[b, ..., 0].item() == 1                                  # :638-639
            prompts.append(self.pos_dynamic_prompt if positive else self.neg_dynamic_prompt)
            for roi_idx in self.roi_indices:                                             # :641-644
                entry = roi_pred_dicts[b][self.roi_ind_names_dict[roi_idx]]
                where = sample_roi_mask[b] == roi_idx
                suvr[b][where] = float(np.nan_to_num(entry["loc"]))
                saliency[b][where] = float(np.nan_to_num(entry["std"]))
        background = x < 1e-4                                                            # :646-647
        suvr = torch.where(background, torch.zeros_like(suvr), suvr)
        saliency = torch.where(background, torch.zeros_like(saliency), saliency)
        prompt = torch.vstack(prompts).to(x.device)
        general = torch.vstack([self.general_dynamic_prompt] * x.size(0)).to(x.device)
        modulated = general + self.deep_modulator_3c(torch.cat((prompt, saliency, suvr), dim=1))   # :651
        fused = self.fusion_layer(torch.cat((modulated, out), dim=1))
        return self.final_act(self.final_pred_head(torch.cat((out, fused), dim=1)))                 # :654-656

    def forward(self, x, covariate=None, roi_pred_dicts=None, sample_roi_mask=None):
        if covariate is not None and x.device != covariate.device:
            covariate = covariate.to(x.device)
        super().forward(x, covariate)                    # :664 duplicate pass; only BN running stats keep its effect
        out, encoder_extractions, _ = super().forward(x, covariate)   # :666
        out = self.forward_modulator_with_uq(x, out, covariate, roi_pred_dicts, sample_roi_mask)
        if not self.training and not self.embeddings_out:
            return out
        projected = [self.projection_heads[i](encoder_extractions[i]) for i in range(self.depth)]
        final_proj = self.final_projection_head(out)
        if self.embeddings_out:
            return out, projected, final_proj, encoder_extractions
        if self.decoder_ds:
            return out, projected, final_proj, []
        return out, projected, final_proj
