"""Drop-in SlowFast video classifier on the dp_b200 kernels (BASELINE config 3).

Mirrors /root/reference/src/models/slowfast.py (SlowNet :11-37, FastNet :43-86, SlowFastEncoder :92-142,
SlowFastClassifier :144-163, SlowFast :165-196) and the 3D-ResNet pieces it uses from src/models/resnet.py
(Swish :63-81, Bottleneck3D :121-200, ResNet3D :202-273): same class names, constructor signatures, module tree
(so `state_dict()` keys match reference checkpoints), construction order (a torch seed gives the reference's
initial weights) and the constructor's shape probe side effect on BatchNorm buffers.

What runs where: every Conv3d (+ BatchNorm3d + ReLU, + fused residual add) -- 74 of them for layers [1,2,2,1] -- is
the tcgen05 / CUDA-core conv family behind the C ABI, including the 3-channel stems (packed-rows fast path) and the
stride-4 temporal lateral convs.  The glue between them (max-pool, squeeze-excite scaling with its two tiny fully
connected layers on (B, C) vectors, temporal sub-sampling of the input clip, MLP head) is PyTorch; every pass over
an activation tensor -- max-pool, squeeze-excite scaling fused with Swish, channel concatenation of the laterals,
global pool -- is a kernel of csrc/slowfast_ops.cu / layout_pool.cu behind the C ABI, forward and backward.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as nn_init

from . import _lib as L
from . import functional as Fn


def _t3(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)



_COUNTED = [False]   # True while SlowFastEncoder.forward has already counted this batch for every BatchNorm3d

class _ConvBN:
    """Adapter handing one (nn.Conv3d, nn.BatchNorm3d) pair of a reference-shaped module tree to the fused layer
    functions.  Not an nn.Module: the parameters stay registered under their reference names."""

    def __init__(self, owner: nn.Module, conv: nn.Conv3d, bn: Optional[nn.BatchNorm3d], slope: float):
        self._owner, self.conv, self.bn = owner, conv, bn
        self._cfg = Fn.LayerCfg(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, slope,
                                bn.eps if bn is not None else 1e-5,
                                bn.momentum if (bn is not None and bn.momentum is not None) else 0.1)
        self._packed, self._packed_stem = Fn.PackedWeights(), Fn.PackedWeights()
        if conv.dilation != (1, 1, 1) or conv.groups != 1:
            raise NotImplementedError("dp_b200: dilated / grouped Conv3d is not on the SlowFast path")

    @property
    def training(self):
        return self._owner.training

    def _bias_fixups(self, before: bool):
        """A conv bias in front of a BatchNorm does not change the normalised output; it only shifts the channel mean.
        Train mode: add momentum*bias to running_mean after the kernel's update.  Eval mode: present running_mean - bias."""
        b = self.conv.bias
        if b is None or self.bn is None or not self.bn.track_running_stats:
            return
        with torch.no_grad():
            if self.training:
                if not before:
                    self.bn.running_mean.add_(b.detach() * self._cfg.momentum)
            else:
                self.bn.running_mean.sub_(b.detach()) if before else self.bn.running_mean.add_(b.detach())

    def __call__(self, x, residual=None, slope_res: float = 1.0, stem_ok: bool = False):
        """x: internal tensor (or, for a stem, the caller's NCDHW clip when `stem_ok`)."""
        bn = self.bn
        self._bias_fixups(before=True)
        try:
            if stem_ok and not Fn.is_internal(x):
                geom = None
                if x.is_cuda and x.dtype == torch.float32 and not x.requires_grad:
                    geom = Fn.stem_geom(self._cfg, x.shape[0], x.shape[2], x.shape[3], x.shape[4])
                if geom is not None:
                    xp = Fn.stem_pack_input(x, geom)
                    z = Fn.stem_conv_bn_act(xp, self.conv.weight, bn.weight, bn.bias, self, geom)
                else:
                    z = Fn.conv_bn_act(Fn.to_internal(x), self.conv.weight, bn.weight, bn.bias, self)
            elif residual is not None:
                z = Fn.conv_bn_act_res(x, self.conv.weight, bn.weight, bn.bias, residual, self, float(slope_res))
            else:
                z = Fn.conv_bn_act(x, self.conv.weight, bn.weight, bn.bias, self)
        finally:
            self._bias_fixups(before=False)
        if self.training and bn.track_running_stats and not _COUNTED[0]:
            bn.num_batches_tracked.add_(1)
        return Fn.tag(z, self._cfg.K)


class Swish(nn.Module):
    """x * sigmoid(x) (reference resnet.py:63-81 hand-writes the backward; csrc/slowfast_ops.cu does too)."""

    def forward(self, x):
        if Fn.is_internal(x):
            return Fn.se_swish(x, None)
        return x * torch.sigmoid(x)


class Bottleneck3D(nn.Module):
    expansion = 4

    def __init__(self, in_planes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None,
                 bias: bool = False, head_conv: int = 1, base_bn_splits: Optional[int] = None, index: int = 0):
        super().__init__()
        if base_bn_splits is not None:
            raise NotImplementedError("dp_b200: SubBatchNorm3d (base_bn_splits) is not used by SlowFast and not provided")
        self.index = index
        if head_conv == 1:
            self.conv1 = nn.Conv3d(in_planes, planes, kernel_size=1, bias=False)
        elif head_conv == 3:
            self.conv1 = nn.Conv3d(in_planes, planes, kernel_size=(3, 1, 1), bias=False, padding=(1, 0, 0))
        else:
            raise ValueError("Unsupported head_conv!")
        self.bn1 = nn.BatchNorm3d(planes)
        self.conv2 = nn.Conv3d(planes, planes, kernel_size=(1, 3, 3), stride=(1, stride, stride), padding=(0, 1, 1), bias=bias)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3 = nn.Conv3d(planes, planes * 4, kernel_size=1, bias=bias)
        self.bn3 = nn.BatchNorm3d(planes * 4)
        self.swish = Swish()
        self.relu = nn.ReLU(inplace=True)
        if self.index % 2 == 0:
            width = self.round_width(planes)
            self.global_pool = nn.AdaptiveAvgPool3d((1, 1, 1))
            self.fc1 = nn.Conv3d(planes, width, kernel_size=1, stride=1)
            self.fc2 = nn.Conv3d(width, planes, kernel_size=1, stride=1)
            self.sigmoid = nn.Sigmoid()
        self.downsample = downsample
        self.stride = stride
        self._l1 = _ConvBN(self, self.conv1, self.bn1, 0.0)
        self._l2 = _ConvBN(self, self.conv2, self.bn2, 0.0)
        self._l3 = _ConvBN(self, self.conv3, self.bn3, 1.0)        # no activation before the residual add
        self._ds = _ConvBN(self, downsample[0], downsample[1], 1.0) if downsample is not None else None

    def round_width(self, width: int, multiplier=0.0625, min_width=8, divisor=8):
        if not multiplier:
            return width
        width *= multiplier
        min_width = min_width or divisor
        width_out = max(min_width, int(width + divisor / 2) // divisor * divisor)
        if width_out < 0.9 * width:
            width_out += divisor
        return int(width_out)

    def forward(self, x):
        x, was_internal = (x, True) if Fn.is_internal(x) else (Fn.to_internal(x), False)
        out = self._l2(self._l1(x))
        if self.index % 2 == 0:
            # squeeze-excite: global mean (dp_avgpool_fwd) -> fc1 -> ReLU -> fc2 -> sigmoid on the (B, C) vector (two
            # tiny fully connected layers, like the MLP head), then channel scale + Swish in ONE pass (dp_se_swish_fwd)
            se = Fn.global_avgpool(out)
            se = F.relu(F.linear(se, self.fc1.weight.view(self.fc1.out_channels, -1), self.fc1.bias))
            se = torch.sigmoid(F.linear(se, self.fc2.weight.view(self.fc2.out_channels, -1), self.fc2.bias))
            out = Fn.se_swish(out, se)
        else:
            out = self.swish(out)
        residual = self._ds(x) if self._ds is not None else x
        out = self._l3(out, residual=residual, slope_res=0.0)   # bn3(conv3) + residual -> ReLU, one fused layer
        return out if was_internal else Fn.to_ncdhw(out)


class ResNet3D(nn.Module):
    def __init__(self, block, layers, **kwargs):
        super().__init__()
        in_channels = kwargs["in_channels"]
        self.alpha = kwargs["alpha"]
        self.slow = kwargs["slow"]
        m = 16
        self.inplanes = (m + m // self.alpha) if self.slow else m // self.alpha
        self.base_bn_splits = kwargs["base_bn_splits"]
        out_channels = m // (1 if self.slow else self.alpha)
        self.layer0 = nn.Sequential(
            nn.Conv3d(in_channels, out_channels, kernel_size=(1, 7, 7), stride=(1, 2, 2), padding=(0, 3, 3)),
            nn.BatchNorm3d(out_channels),
            nn.ReLU(inplace=True),
            nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)),
        )
        q = 1 if self.slow else self.alpha
        self.layer1 = self._make_layer(block, m // q, layers[0], head_conv=1 if self.slow else 3,
                                       base_bn_splits=self.base_bn_splits)
        self.layer2 = self._make_layer(block, 2 * m // q, layers[1], stride=2, head_conv=1 if self.slow else 3,
                                       base_bn_splits=self.base_bn_splits)
        self.layer3 = self._make_layer(block, 4 * m // q, layers[2], stride=2, head_conv=3,
                                       base_bn_splits=self.base_bn_splits)
        self.layer4 = self._make_layer(block, 8 * m // q, layers[3], stride=2, head_conv=3,
                                       base_bn_splits=self.base_bn_splits)
        self._stem = _ConvBN(self, self.layer0[0], self.layer0[1], 0.0)

    def init_params(self):
        for mod in self.modules():
            if isinstance(mod, nn.Conv3d):
                nn_init.xavier_normal_(mod.weight)
            elif isinstance(mod, nn.BatchNorm3d) and mod.weight is not None:
                nn_init.constant_(mod.weight, 1)

    def forward(self, x):
        raise NotImplementedError("use each pathway network's forward function")

    def _make_layer(self, block, planes: int, blocks: int = 3, stride: int = 1, head_conv: int = 1,
                    base_bn_splits: Optional[int] = None):
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv3d(self.inplanes, planes * block.expansion, kernel_size=1, stride=(1, stride, stride), bias=False),
                nn.BatchNorm3d(planes * block.expansion),
            )
        else:
            downsample = None
        layers = [block(self.inplanes, planes, stride, downsample, head_conv=head_conv, base_bn_splits=base_bn_splits)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, head_conv=head_conv, base_bn_splits=base_bn_splits))
        self.inplanes += self.slow * block.expansion * planes // self.alpha
        return nn.Sequential(*layers)

    # ---- shared pieces of the two pathways ----
    def _layer0(self, x):
        """conv(+bias) -> BN -> ReLU on the conv kernels (packed-rows stem for 3-channel clips), then the max-pool."""
        z = self._stem(x, stem_ok=True)
        mp = self.layer0[3]
        if (_t3(mp.kernel_size), _t3(mp.stride), _t3(mp.padding)) != ((1, 3, 3), (1, 2, 2), (0, 1, 1)) or mp.ceil_mode:
            raise NotImplementedError("dp_b200: only the reference's MaxPool3d((1,3,3),(1,2,2),(0,1,1)) is provided")
        return Fn.maxpool_hw(z)

    @staticmethod
    def _pooled(x):
        return Fn.global_avgpool(x)


def _cat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return Fn.concat_channels(a, b)


class SlowNet(ResNet3D):
    def __init__(self, blocks, layers, **kwargs):
        super().__init__(blocks, layers, **kwargs)
        self.init_params()

    def forward(self, x: Tuple[torch.Tensor, List[torch.Tensor]]):
        x, laterals = x
        x = self._layer0(x)
        for layer, lat in zip((self.layer1, self.layer2, self.layer3, self.layer4), laterals):
            x = layer(_cat_channels(x, lat))
        return self._pooled(x)


def resnet50_s(block=Bottleneck3D, layers=[3, 4, 6, 3], **kwargs):
    return SlowNet(block, layers, **kwargs)


class _Lateral:
    def __init__(self, conv: nn.Conv3d):
        self.conv = conv
        self._cfg = Fn.LayerCfg(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, 1.0)
        self._packed = Fn.PackedWeights()

    def __call__(self, x):
        return Fn.tag(Fn.ConvFn.apply(x, self.conv.weight, self), self._cfg.K)


class FastNet(ResNet3D):
    def __init__(self, blocks, layers, **kwargs):
        super().__init__(blocks, layers, **kwargs)
        alpha = kwargs["alpha"]
        k, s, p = (alpha + 2, 1, 1), (alpha, 1, 1), (1, 0, 0)
        m = 16
        self.l_maxpool = nn.Conv3d(m // self.alpha, m // self.alpha, kernel_size=k, stride=s, bias=False, padding=p)
        self.l_layer1 = nn.Conv3d(4 * m // self.alpha, 4 * m // self.alpha, kernel_size=k, stride=s, bias=False, padding=p)
        self.l_layer2 = nn.Conv3d(8 * m // self.alpha, 8 * m // self.alpha, kernel_size=k, stride=s, bias=False, padding=p)
        self.l_layer3 = nn.Conv3d(16 * m // self.alpha, 16 * m // self.alpha, kernel_size=k, stride=s, bias=False, padding=p)
        self.init_params()
        self._lats = [_Lateral(c) for c in (self.l_maxpool, self.l_layer1, self.l_layer2, self.l_layer3)]

    def forward(self, x: torch.Tensor):
        laterals = []
        x = self._layer0(x)
        laterals.append(self._lats[0](x))
        for i, layer in enumerate((self.layer1, self.layer2, self.layer3)):
            x = layer(x)
            laterals.append(self._lats[i + 1](x))
        x = self.layer4(x)
        return self._pooled(x), laterals


def resnet50_f(block=Bottleneck3D, layers=[3, 4, 6, 3], **kwargs):
    return FastNet(block, layers, **kwargs)


class SlowFastEncoder(nn.Module):
    def __init__(self, input_shape: Tuple[int, int, int, int] = (3, 8, 112, 112), block=Bottleneck3D,
                 layers: List[int] = [3, 4, 6, 3], alpha: int = 4, tau_fast: int = 1):
        super().__init__()
        self.input_shape = input_shape
        self.seq_len = input_shape[1]
        self.in_channels = input_shape[0]
        self.alpha = alpha
        self.tau_fast = tau_fast
        self.slownet = resnet50_s(block=block, layers=layers, alpha=alpha, in_channels=self.in_channels, slow=1,
                                  base_bn_splits=None)
        self.fastnet = resnet50_f(block=block, layers=layers, alpha=alpha, in_channels=self.in_channels, slow=0,
                                  base_bn_splits=None)
        self._out_dim = (8 * 16 * block.expansion) + (8 * 16 // alpha) * block.expansion

    def split_slow_fast(self, x: torch.Tensor):
        tau_slow = self.tau_fast * self.alpha
        return x[:, :, ::tau_slow, :, :], x[:, :, ::self.tau_fast, :, :]

    def forward(self, x: torch.Tensor):
        L.require_device()
        # one multi-tensor increment of every BatchNorm3d batch counter (the per-layer increments are 4-us launches
        # between the convs); the layers see _COUNTED and skip theirs
        counted = self.training and not _COUNTED[0]
        if counted:
            bns = [m for m in self.modules() if isinstance(m, nn.BatchNorm3d) and m.track_running_stats]
            if bns:
                torch._foreach_add_([m.num_batches_tracked for m in bns], 1)
            _COUNTED[0] = True
        try:
            x_slow, x_fast = self.split_slow_fast(x)
            x_fast, laterals = self.fastnet(x_fast.contiguous())
            x_slow = self.slownet((x_slow.contiguous(), laterals))
            return torch.cat([x_slow, x_fast], dim=1)
        finally:
            if counted:
                _COUNTED[0] = False

    def get_output_shape(self):
        """The reference pushes a zero clip through both pathways on the CPU in TRAINING mode (slowfast.py:137-141).
        The width is known (2048/expansion... = slow 512 + fast 512/alpha channels); the probe's side effect on the
        buffers is reproduced: every BatchNorm3d has seen one all-zero batch (running_var <- 0.9, one batch tracked),
        and the stems, whose conv has a bias, have seen a constant batch equal to that bias."""
        with torch.no_grad():
            for net in (self.slownet, self.fastnet):
                for mod in net.modules():
                    if isinstance(mod, nn.BatchNorm3d) and mod.track_running_stats and mod.momentum is not None:
                        mod.running_mean.mul_(1.0 - mod.momentum)
                        mod.running_var.mul_(1.0 - mod.momentum)
                        mod.num_batches_tracked.add_(1)
                net.layer0[1].running_mean.add_(net.layer0[0].bias.detach() * net.layer0[1].momentum)
        return torch.Size([1, self._out_dim])


class SlowFastClassifier(nn.Module):
    def __init__(self, input_dim: int, num_classes: int = 2, alpha: float = 1.0):
        super().__init__()
        self.input_dim = input_dim
        self.classifier = nn.Sequential(nn.Linear(input_dim, input_dim // 2), nn.BatchNorm1d(input_dim // 2), nn.ELU(alpha),
                                        nn.Linear(input_dim // 2, num_classes))

    def forward(self, x: torch.Tensor):
        return self.classifier(x)


class SlowFast(nn.Module):
    def __init__(self, input_shape: Tuple[int, int, int, int] = (3, 8, 112, 112), block=Bottleneck3D,
                 layers: List[int] = [3, 4, 6, 3], alpha: int = 4, tau_fast: int = 1, num_classes: int = 2,
                 alpha_elu: float = 1.0):
        super().__init__()
        self.input_shape = input_shape
        self.encoder = SlowFastEncoder(input_shape, block, layers, alpha, tau_fast)
        cls_input_dim = self.encoder.get_output_shape()[-1]
        self.classifier = SlowFastClassifier(cls_input_dim, num_classes, alpha_elu)

    def encode(self, x: torch.Tensor):
        with torch.no_grad():
            x = self.encoder.forward(x)
            return x.view(x.size(0), -1)

    def forward(self, x: torch.Tensor):
        return self.classifier.forward(self.encoder.forward(x))

    def summary(self, device: str = "cpu", show_input: bool = True, show_hierarchical: bool = True,
                print_summary: bool = False, show_parent_layers: bool = True):
        n = sum(p.numel() for p in self.parameters())
        print(f"{self}\ninput (B,{','.join(map(str, self.input_shape))}) | parameters: {n:,}")
