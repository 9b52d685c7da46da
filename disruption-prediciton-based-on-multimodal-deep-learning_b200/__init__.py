"""B200-native (sm_100a) hot path of ZINZINBIN/Disruption-Prediciton-based-on-Multimodal-Deep-Learning:
the R(2+1)D SpatioTemporalConv stack, Focal / LDAM / CE losses and RS / RW / DRW re-weighting, behind the
reference's own nn.Module and loss signatures.  Import through the `dp_b200` alias at the repo root:

    from dp_b200.R2Plus1D import R2Plus1DClassifier
    from dp_b200.loss import FocalLoss, LDAMLoss, CELoss
"""
from . import _lib
from . import functional
from . import R2Plus1D
from . import loss
from . import optim
from . import distributed
from . import inference
from . import graph
from . import MultiModal
from . import slowfast
from .functional import compute_mode, get_compute_mode, set_compute_mode, set_conv_impl
from .loss import CELoss, FocalLoss, ImbalancedDatasetSampler, LDAMLoss, drw_betas, drw_class_weights, rw_class_weights
from .R2Plus1D import (Conv3dBlock, R2Plus1DClassifier, R2Plus1DNet, SpatioTemporalConv, SpatioTemporalResBlock,
                       SpatioTemporalResLayer)

__all__ = [
    "R2Plus1DClassifier", "R2Plus1DNet", "SpatioTemporalConv", "SpatioTemporalResBlock", "SpatioTemporalResLayer",
    "Conv3dBlock", "FocalLoss", "LDAMLoss", "CELoss", "ImbalancedDatasetSampler", "rw_class_weights",
    "drw_class_weights", "drw_betas", "compute_mode", "set_compute_mode", "get_compute_mode", "set_conv_impl",
]
