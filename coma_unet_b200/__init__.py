"""coma_unet_b200: B200-native forward/backward of CoMA-UNet's covariate-modulated attention U-Net.

Same public surface as the reference for the hot path (attn_unet_data_parallel.py:120-693,
criterions.py:124-211,485-644): import the classes from here instead of the reference modules.
Requires the in-tree CUDA library (coma_unet_b200/csrc/libcoma_b200.so); there is no CPU fallback.
"""
from . import _lib
from . import metrics
from .criterions import GenerativeContrastiveLoss, RnCLoss, RoiMSE
from .data import DevicePrefetcher, HostSink, SyntheticVolumeDataset, prepare_geometry, prepare_volumes
from .model import (AttentionLayer, ContrastiveAttentionUNET_DP, ObservableAttentionBlock, ObservableAttentionUnet,
                    ProjectionHead, StackedFusionConvLayers, UpBlock)
from .graph import GraphedInference, GraphedTrainStep
from .parallel import DataParallelEngine
from .train import train_dp

__all__ = ["ContrastiveAttentionUNET_DP", "ObservableAttentionUnet", "AttentionLayer", "ObservableAttentionBlock",
           "UpBlock", "ProjectionHead", "StackedFusionConvLayers", "RoiMSE", "RnCLoss", "GenerativeContrastiveLoss",
           "SyntheticVolumeDataset", "DevicePrefetcher", "HostSink", "DataParallelEngine", "train_dp", "metrics", "_lib",
           "prepare_geometry", "prepare_volumes", "GraphedTrainStep", "GraphedInference"]
