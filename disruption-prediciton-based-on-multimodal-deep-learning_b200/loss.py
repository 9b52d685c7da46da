"""Drop-in CE / Focal / LDAM losses and RW / DRW / RS re-weighting on the dp_b200 kernels.

Mirrors /root/reference/src/loss.py (FocalLoss :14-34, LDAMLoss :37-69, CELoss :71-81): same class
names, constructor signatures, `.forward(logits, target)`, `.update_weight`, `.update_m_list`,
`.model_type`, `.m_list`, `.weight`.  Each forward is ONE fused CUDA launch that also produces the
logit gradients (csrc/loss.cu); the LDAM margins stay on the device (the reference multiplies them on
the host every step, loss.py:62-64).

Reductions follow the reference exactly: CE and Focal are SUMS over the batch, LDAM is the
weighted MEAN  sum_i w[y_i] CE_i / sum_i w[y_i].
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import functional as Fn


class _DeviceCache:
    """Device-resident fp32 copy of a small host/device tensor, refreshed when the source changes."""

    def __init__(self):
        self.key = None
        self.val = None

    def get(self, t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
        if t is None:
            return None
        key = (id(t), t._version, str(device))
        if key != self.key:
            self.val = t.detach().to(device=device, dtype=torch.float32).contiguous()
            self.key = key
        return self.val


class FocalLoss(nn.Module):
    """sum_i w[y_i] * (1 - exp(-CE_i))**gamma * CE_i   (reference loss.py:14-34; CE unweighted, sum)."""

    def __init__(self, weight: Optional[torch.Tensor] = None, gamma: float = 2.0):
        super().__init__()
        assert gamma >= 0, "gamma should be positive"
        self.model_type = "Focal"
        self.gamma = gamma
        self.weight = weight
        self._w = _DeviceCache()

    def update_weight(self, weight: Optional[torch.Tensor] = None):
        self.weight = weight

    def forward(self, input: torch.Tensor, target: torch.Tensor):
        w = self._w.get(self.weight, input.device)
        return Fn.LossFn.apply(input, target, w, None, L.LOSS_FOCAL, float(self.gamma), 1.0)


class LDAMLoss(nn.Module):
    """Label-distribution-aware margin loss (reference loss.py:37-69)."""

    def __init__(self, cls_num_list: Optional[List], max_m: float = 0.5, weight: Optional[torch.Tensor] = None,
                 s: int = 30):
        super().__init__()
        assert s > 0, "s should be positive"
        self.model_type = "LDAM"
        self.s = s
        self.max_m = max_m
        self.weight = weight
        self._w = _DeviceCache()
        self._m = _DeviceCache()
        if cls_num_list:
            self.update_m_list(cls_num_list)

    def update_weight(self, weight: Optional[torch.Tensor] = None):
        self.weight = weight

    def update_m_list(self, cls_num_list: List):
        # m_c = max_m * n_c^(-1/4) / max_j n_j^(-1/4), in float64 then rounded to float32 (loss.py:52-56)
        m_list = 1.0 / np.sqrt(np.sqrt(cls_num_list))
        m_list = m_list * (self.max_m / np.max(m_list))
        self.m_list = torch.FloatTensor(m_list)

    def forward(self, x: torch.Tensor, target: torch.Tensor):
        w = self._w.get(self.weight, x.device)
        m = self._m.get(self.m_list, x.device)
        if m.numel() != x.shape[1]:
            raise L.DpError(f"LDAM: {m.numel()} margins for {x.shape[1]} classes")
        return Fn.LossFn.apply(x, target, w, m, L.LOSS_LDAM, 0.0, float(self.s))


class CELoss(nn.Module):
    """Weighted cross entropy, SUM over the batch (reference loss.py:71-81)."""

    def __init__(self, weight: Optional[torch.Tensor] = None):
        super().__init__()
        self.model_type = "CE"
        self.weight = weight
        self._w = _DeviceCache()

    def update_weight(self, weight: Optional[torch.Tensor] = None):
        self.weight = weight

    def forward(self, x: torch.Tensor, target: torch.Tensor):
        w = self._w.get(self.weight, x.device)
        return Fn.LossFn.apply(x, target, w, None, L.LOSS_CE, 0.0, 1.0)


# ---------------------------------------------------------------------------------------------
# class re-weighting (host side, two floats per epoch)
# ---------------------------------------------------------------------------------------------
def rw_class_weights(cls_num_list, use_weighting: bool = True) -> torch.Tensor:
    """RW: inverse-frequency weights normalised to sum 1, else [1, 1]
    (reference train_vision_network.py:312-318)."""
    if use_weighting:
        w = 1.0 / np.array(cls_num_list)
        w = w / np.sum(w)
    else:
        w = np.array([1, 1])
    return torch.FloatTensor(w)


def drw_betas(beta: float) -> List[float]:
    """Vision trainer's DRW schedule [0, b, 2b, 3b] (reference train_vision_network.py:322-335)."""
    return [0, beta, beta * 2, beta * 3]


def drw_class_weights(epoch: int, num_epoch: int, betas: List[float], cls_num_list) -> torch.Tensor:
    """DRW: effective-number class weights for this epoch, normalised to sum = #classes
    (reference src/train.py:318-329)."""
    idx = epoch // int(num_epoch / len(betas))
    if idx >= len(betas):
        idx = len(betas) - 1
    beta = betas[idx]
    effective_num = 1.0 - np.power(beta, cls_num_list)
    per_cls_weights = (1.0 - beta) / np.array(effective_num)
    per_cls_weights = per_cls_weights / np.sum(per_cls_weights) * len(cls_num_list)
    return torch.FloatTensor(per_cls_weights)


class ImbalancedDatasetSampler(torch.utils.data.sampler.Sampler):
    """RS: multinomial re-sampling with 1/class-count weights (reference src/utils/sampler.py:5-35).
    Host-side, once per epoch; any object with `.labels` works as `dataset`."""

    def __init__(self, dataset, indices=None, num_samples=None, callback_get_label=None):
        self.indices = list(range(len(dataset))) if indices is None else indices
        self.callback_get_label = callback_get_label
        self.num_samples = len(self.indices) if num_samples is None else num_samples
        counts = {}
        labels = [self._get_label(dataset, i) for i in self.indices]
        for lab in labels:
            counts[lab] = counts.get(lab, 0) + 1
        self.weights = torch.DoubleTensor([1.0 / counts[lab] for lab in labels])

    def _get_label(self, dataset, idx):
        return dataset.labels[idx]

    def __iter__(self):
        return (self.indices[i] for i in torch.multinomial(self.weights, self.num_samples, replacement=True))

    def __len__(self):
        return self.num_samples
