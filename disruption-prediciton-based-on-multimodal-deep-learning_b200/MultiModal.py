"""R(2+1)D-backed multimodal fusion (BASELINE config 4).

The reference's multimodal models (/root/reference/src/models/MultiModal.py:10-163) hard-wire ViViT as the video
encoder, and its 3-head `MultiModalModel_GB` cannot even be constructed (`Transformer` has no `feature_dims`,
SURVEY D3).  There is therefore no reference model to be a drop-in FOR; this module defines the model config 4
names, keeping the reference's structure wherever it exists:

  * video branch  = `R2Plus1DNet` on the dp_b200 kernels (128-d feature), head = the R2Plus1DClassifier MLP
  * 0D branch     = `TransformerEncoder` / `Transformer` with the reference's layer structure and attribute names
                    (src/models/transformer.py:39-138): Conv1d x2 -> BatchNorm1d -> ReLU, sinusoidal positions,
                    causal nn.TransformerEncoder, mean over time, Linear -> LayerNorm -> GELU.  Plain PyTorch: it
                    is 0.1 % of the FLOPs and not on the hot path
  * fusion        = concat -> `connector` -> `classifier`, exactly MultiModal.py:21-31
  * loss          = `GradientBlending` (src/GradientBlending.py:20-50): w_v L(vis) + w_t L(ts) + w_m L(fusion), each L
                    one of the fused dp_b200 losses

Parity is part-wise (tests/test_gpu_multimodal.py): the video feature against the oracle encoder, the blended
loss against the oracle loss formulas, the fusion head against the same torch layers.
"""
from __future__ import annotations

import math
from typing import Dict, List, Literal, Optional

import torch
import torch.nn as nn

from .R2Plus1D import R2Plus1DNet


class NoiseLayer(nn.Module):
    """Additive Gaussian input noise in training mode (reference src/models/NoiseLayer.py)."""

    def __init__(self, mean: float = 0.0, std: float = 1e-2):
        super().__init__()
        self.mean, self.std = mean, std

    def forward(self, x: torch.Tensor):
        if not self.training:
            return x
        return x + self.mean + torch.randn_like(x) * self.std


class PositionalEncoding(nn.Module):
    """Fixed sinusoidal table, shape (max_len, 1, d_model) (reference transformer.py:10-33)."""

    def __init__(self, d_model: int, max_len: int = 128):
        super().__init__()
        pos = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, d_model)
        pe[:, 0::2] = torch.sin(pos * freq)
        pe[:, 1::2] = torch.cos(pos * freq)[:, : d_model // 2]
        self.register_buffer("pe", pe.unsqueeze(1))

    def forward(self, x: torch.Tensor):     # (seq, batch, d_model)
        return x + self.pe[: x.size(0)]


class GELU(nn.Module):
    """tanh approximation, as the reference writes it out (transformer.py:35-37)."""

    def forward(self, x):
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


class TransformerEncoder(nn.Module):
    def __init__(self, n_features: int = 11, kernel_size: int = 3, feature_dims: int = 256, max_len: int = 128,
                 n_layers: int = 1, n_heads: int = 8, dim_feedforward: int = 1024, dropout: float = 0.1):
        super().__init__()
        self.n_features, self.max_len, self.feature_dims = n_features, max_len, feature_dims
        self.src_mask = None
        self.noise = NoiseLayer(mean=0, std=1e-3)
        pad = (kernel_size - 1) // 2
        self.filter = nn.Sequential(
            nn.Conv1d(n_features, feature_dims, kernel_size, 1, pad),
            nn.Conv1d(feature_dims, feature_dims, kernel_size, 1, pad),
            nn.BatchNorm1d(feature_dims),
            nn.ReLU(),
        )
        self.pos_enc = PositionalEncoding(feature_dims, max_len)
        layer = nn.TransformerEncoderLayer(d_model=feature_dims, nhead=n_heads, dropout=dropout,
                                           dim_feedforward=dim_feedforward, activation=GELU())
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=n_layers)
        self.connector = nn.Sequential(nn.Linear(feature_dims, feature_dims), nn.LayerNorm(feature_dims), nn.GELU())

    def forward(self, x: torch.Tensor):     # (batch, seq, n_features)
        x = self.filter(self.noise(x).permute(0, 2, 1)).permute(2, 0, 1)          # (seq, batch, d)
        n = x.size(0)
        self.src_mask = torch.full((n, n), float("-inf"), device=x.device).triu(1)  # causal
        # the mask IS the causal mask (reference transformer.py:105,111-114); saying so spares nn.TransformerEncoder its
        # mask inspection, a host synchronisation per call that also forbids CUDA-graph capture of the step
        x = self.transformer_encoder(self.pos_enc(x), self.src_mask, is_causal=True).mean(dim=0)
        return self.connector(x)


class Transformer(nn.Module):
    def __init__(self, n_features: int = 11, kernel_size: int = 5, feature_dims: int = 256, max_len: int = 128,
                 n_layers: int = 1, n_heads: int = 8, dim_feedforward: int = 1024, dropout: float = 0.1,
                 cls_dims: int = 128, n_classes: int = 2):
        super().__init__()
        self.max_len, self.n_features, self.feature_dims = max_len, n_features, feature_dims
        self.encoder = TransformerEncoder(n_features, kernel_size, feature_dims, max_len, n_layers, n_heads,
                                          dim_feedforward, dropout)
        self.classifier = nn.Sequential(nn.Linear(feature_dims, cls_dims), nn.LayerNorm(cls_dims), GELU(),
                                        nn.Linear(cls_dims, n_classes))

    def encode(self, x: torch.Tensor):
        with torch.no_grad():
            return self.encoder(x)

    def forward(self, x: torch.Tensor):
        return self.classifier(self.encoder(x))


def _fusion_heads(dims: int, n_classes: int):
    connector = nn.Sequential(nn.Linear(dims, dims // 2), nn.ReLU())
    classifier = nn.Sequential(nn.Linear(dims // 2, dims // 2), nn.LayerNorm(dims // 2), nn.ReLU(),
                               nn.Linear(dims // 2, n_classes))
    return connector, classifier


class MultiModalR2Plus1D(nn.Module):
    """`MultiModalModel` (reference MultiModal.py:10-48) with the R(2+1)D encoder as `encoder_video`.
    args_video: {"layer_sizes": [...], "alpha": float}; args_0D: TransformerEncoder kwargs."""

    def __init__(self, n_classes: int, args_video: Dict, args_0D: Dict):
        super().__init__()
        self.n_classes, self.args_video, self.args_0D = n_classes, args_video, args_0D
        self.encoder_video = R2Plus1DNet(args_video.get("layer_sizes", [1, 2, 2, 1]), alpha=args_video.get("alpha", 1.0))
        self.encoder_0D = TransformerEncoder(**args_0D)
        dims = self.encoder_0D.feature_dims + self.encoder_video.out_features
        self.connector, self.classifier = _fusion_heads(dims, n_classes)

    def forward(self, x_video: torch.Tensor, x_0D: torch.Tensor):
        h = torch.cat([self.encoder_video(x_video), self.encoder_0D(x_0D)], dim=1)
        return self.classifier(self.connector(h))

    def encode(self, x_vis: torch.Tensor, x_0D: torch.Tensor):
        with torch.no_grad():
            h_vis, h_0D = self.encoder_video(x_vis), self.encoder_0D(x_0D)
            return self.connector(torch.cat([h_vis, h_0D], dim=1)), h_vis, h_0D


class MultiModalR2Plus1D_GB(nn.Module):
    """Three-head variant for Gradient Blending (what MultiModal.py:56-163 intends): uni-modal logits from each
    branch's own head plus fusion logits from the concatenated latents.  `use_stream` as in the reference."""

    def __init__(self, n_classes: int, args_video: Dict, args_0D: Dict,
                 use_stream: Literal["video", "0D", "multi", "multi-GB"] = "multi-GB"):
        super().__init__()
        self.n_classes, self.args_video, self.args_0D, self.use_stream = n_classes, args_video, args_0D, use_stream
        alpha = args_video.get("alpha", 1.0)
        self.vis_model = nn.Module()
        self.vis_model.res2plus1d = R2Plus1DNet(args_video.get("layer_sizes", [1, 2, 2, 1]), alpha=alpha)
        d = self.vis_model.res2plus1d.out_features
        self.vis_model.linear = nn.Sequential(nn.Linear(d, d // 2), nn.BatchNorm1d(d // 2), nn.ELU(alpha),
                                              nn.Linear(d // 2, n_classes))      # the R2Plus1DClassifier head
        self.ts_model = Transformer(**args_0D)
        self.connector, self.classifier = _fusion_heads(self.ts_model.feature_dims + d, n_classes)
        self.vis_latent: Optional[torch.Tensor] = None
        self.ts_latent: Optional[torch.Tensor] = None

    def update_use_stream(self, use_stream: str):
        self.use_stream = use_stream

    def forward(self, x_vis: torch.Tensor, x_ts: torch.Tensor):
        if self.use_stream == "video":
            return self.vis_model.linear(self.vis_model.res2plus1d(x_vis))
        if self.use_stream == "0D":
            return self.ts_model(x_ts)
        self.vis_latent = self.vis_model.res2plus1d(x_vis)
        self.ts_latent = self.ts_model.encoder(x_ts)
        out_vis = self.vis_model.linear(self.vis_latent)
        out_ts = self.ts_model.classifier(self.ts_latent)
        out_multi = self.classifier(self.connector(torch.cat([self.vis_latent, self.ts_latent], dim=1)))
        return out_multi if self.use_stream == "multi" else (out_multi, out_vis, out_ts)

    def encode(self, x_vis: torch.Tensor, x_0D: torch.Tensor):
        with torch.no_grad():
            v, t = self.vis_model.res2plus1d(x_vis), self.ts_model.encoder(x_0D)
            return self.connector(torch.cat([v, t], dim=1)), v, t


class GradientBlending(nn.Module):
    """loss = w_vis L_vis(out_vis) + w_ts L_ts(out_ts) + w_multi L_multi(out_multi), each scaled by `loss_scale`
    (reference src/GradientBlending.py:20-50; trainer weights .1/.4/.5 at train_multimodal.py:375-385)."""

    def __init__(self, loss_vis: nn.Module, loss_ts: nn.Module, loss_vis_ts: nn.Module, vis_weight: float = 0.0,
                 ts_weight: float = 0.0, vis_ts_weight: float = 1.0, loss_scale: float = 1.0):
        super().__init__()
        self.loss_vis, self.loss_ts, self.loss_vis_ts = loss_vis, loss_ts, loss_vis_ts
        self.vis_weight, self.ts_weight, self.vis_ts_weight, self.loss_scale = vis_weight, ts_weight, vis_ts_weight, loss_scale

    def update_weights(self, ws: Dict):
        self.vis_weight, self.ts_weight, self.vis_ts_weight = ws["video"], ws["0D"], ws["multi"]

    def forward(self, vis_ts_out: torch.Tensor, vis_out: torch.Tensor, ts_out: torch.Tensor, target: torch.Tensor):
        s = self.loss_scale
        return (self.loss_vis(vis_out, target) * s * self.vis_weight + self.loss_ts(ts_out, target) * s * self.ts_weight +
                self.loss_vis_ts(vis_ts_out, target) * s * self.vis_ts_weight)
