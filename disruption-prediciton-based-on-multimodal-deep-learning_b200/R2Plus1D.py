"""Drop-in R(2+1)D video encoder / classifier running on the dp_b200 CUDA kernels.

Mirrors the public interface of the reference module /root/reference/src/models/R2Plus1D.py:
same class names, constructor signatures and defaults, sub-module names (so `state_dict()` keys and
`load_state_dict` of reference checkpoints match: `res2plus1d.conv1.spatio_conv.conv.weight`, ...),
same construction order (so a given torch seed yields the same initial weights), same
`forward / encode / summary / input_size` surface.  What differs is what runs: every
Conv3d / BatchNorm3d / LeakyReLU / residual add / pool here is a hand-written sm_100a kernel behind
the C ABI in include/dp_b200.h.  There is no cuDNN call and no CPU path; CPU tensors raise.

Interface notes carried over from the reference (file:line there):
  * Conv3dBlock promotes int kernel/stride/padding to (1,k,k)/(1,s,s)/(0,p,p)      (:29-42)
  * SpatioTemporalConv: mid = floor(kt*kh*kw*Cin*Cout / (kh*kw*Cin + kt*Cout)); the stem uses 45  (:126-155)
  * SpatioTemporalResBlock builds its SpatioTemporalConvs WITHOUT passing alpha, so the inner
    activations use the SpatioTemporalConv default slope 0.01; only the block-end LeakyReLU and the
    stem use the model's alpha                                                      (:172-179, :210)
  * head: Linear(128,64) -> BatchNorm1d -> ELU(alpha) -> Linear(64,num_classes)     (:243-248)
"""
from __future__ import annotations

import math
from typing import List, Tuple, Union

import torch
import torch.nn as nn
from torch.nn.modules.utils import _triple

from . import _lib as L
from . import functional as Fn


def _hooked(m: nn.Module) -> bool:
    return bool(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or m._backward_pre_hooks)


def _any_inner_hooks(m: nn.Module) -> bool:
    for sub in m.modules():
        if sub is not m and _hooked(sub):
            return True
    return False


def _enter(x: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    """Accept either a caller-facing NCDHW fp32 tensor or an internal NDHWC tensor."""
    if Fn.is_internal(x):
        return x, True
    return Fn.to_internal(x), False


def _leave(z: torch.Tensor, was_internal: bool) -> torch.Tensor:
    """Internal tensors stay internal between our own modules; outside callers get NCDHW fp32."""
    return z if was_internal else Fn.to_ncdhw(z)


def _call(child: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """Run a child module on an internal tensor.  A child that carries hooks (GradCAM registers forward and
    backward hooks on `res2plus1d.conv5`, reference visualize_cam.py:75-76) is handed -- and hands back --
    caller-facing NCDHW fp32 tensors, so its hooks observe what they would observe in the reference."""
    if _hooked(child):
        out = child(Fn.to_ncdhw(x))
        return _enter(out)[0]
    return child(x)



_COUNTED = [False]   # True while an enclosing R2Plus1DNet.forward has already counted this batch for every BatchNorm3d

class Conv3dBlock(nn.Module):
    """Conv3d -> BatchNorm3d -> LeakyReLU(alpha)   (reference R2Plus1D.py:25-58)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size=3, stride=1, dilation: int = 1, padding=1,
                 bias: bool = False, alpha: float = 0.01):
        super().__init__()
        strides = stride if type(stride) == tuple else (1, stride, stride)
        kernel_sizes = kernel_size if type(kernel_size) == tuple else (1, kernel_size, kernel_size)
        paddings = padding if type(padding) == tuple else (0, padding, padding)
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size=kernel_sizes, stride=strides, padding=paddings,
                              dilation=dilation, bias=bias)
        self.bn = nn.BatchNorm3d(out_channels)
        self.relu = nn.LeakyReLU(alpha)
        if _triple(dilation) != (1, 1, 1):
            raise NotImplementedError("dp_b200: dilation != 1 is not on the R(2+1)D path (reference always passes 1)")
        if bias:
            raise NotImplementedError("dp_b200: Conv3dBlock(bias=True) is not on the R(2+1)D path (all its convs are bias-free)")
        self._cfg = Fn.LayerCfg(in_channels, out_channels, kernel_sizes, strides, paddings, alpha)
        self._packed = Fn.PackedWeights()
        self._packed_stem = Fn.PackedWeights()

    def _refresh_cfg(self):
        c = self._cfg
        c.slope = float(self.relu.negative_slope)
        c.eps = float(self.bn.eps)
        if self.bn.momentum is None:
            raise NotImplementedError("dp_b200: BatchNorm3d(momentum=None) (cumulative average) is not supported")
        c.momentum = float(self.bn.momentum)

    def _stem_geom(self, x: torch.Tensor):
        """Geometry of the packed-rows stem fast path if it applies to this block and this caller-facing clip
        (NCDHW fp32 without grad, or (B,T,H,W,3) uint8 frames), else None."""
        if Fn.is_internal(x) or not x.is_cuda or x.dim() != 5 or x.requires_grad:
            return None
        if x.dtype == torch.uint8:
            B, T, H, W, Cc = x.shape
        elif x.dtype == torch.float32:
            B, Cc, T, H, W = x.shape
        else:
            return None
        if Cc != self._cfg.C:
            return None
        return Fn.stem_geom(self._cfg, B, T, H, W)

    def forward_stem(self, x: torch.Tensor, geom, mean_bgr=None):
        """External clip in, INTERNAL activation out (used by R2Plus1DNet for its first layer)."""
        self._refresh_cfg()
        xp = Fn.stem_pack_input(x, geom, mean_bgr)
        z = Fn.stem_conv_bn_act(xp, self.conv.weight, self.bn.weight, self.bn.bias, self, geom)
        if self.training and self.bn.track_running_stats and not _COUNTED[0]:
            self.bn.num_batches_tracked.add_(1)
        return Fn.tag(z, self._cfg.K)

    def forward(self, x: torch.Tensor):
        x, was_internal = _enter(x)
        self._refresh_cfg()
        z = Fn.conv_bn_act(x, self.conv.weight, self.bn.weight, self.bn.bias, self)
        if self.training and self.bn.track_running_stats and not _COUNTED[0]:
            self.bn.num_batches_tracked.add_(1)
        Fn.tag(z, self._cfg.K)
        return _leave(z, was_internal)


class SpatioTemporalConv(nn.Module):
    """Factorised (2+1)D convolution: spatial Conv3dBlock then temporal Conv3dBlock (reference :115-162)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size=(3, 1, 1), stride=(1, 1, 1), dilation: int = 1,
                 padding=(1, 1, 1), bias: bool = False, alpha: float = 0.01, is_first: bool = False):
        super().__init__()
        if type(kernel_size) == int:
            kernel_size = _triple(kernel_size)
        if type(stride) == int:
            stride = _triple(stride)
        if type(padding) == int:
            padding = _triple(padding)
        kt, kh, kw = kernel_size
        if is_first:
            # stem: the given kernel is already spatial; temporal part fixed to 3 taps, 45 mid channels
            mid = 45
            self.spatio_conv = Conv3dBlock(in_channels, mid, kernel_size, (1, stride[1], stride[2]), dilation, padding,
                                           False, alpha)
            self.temporal_conv = Conv3dBlock(mid, out_channels, (3, 1, 1), (stride[0], 1, 1), dilation, (1, 0, 0),
                                             False, alpha)
        else:
            mid = int(math.floor((kt * kh * kw * in_channels * out_channels) /
                                 (kh * kw * in_channels + kt * out_channels)))
            self.spatio_conv = Conv3dBlock(in_channels, mid, (1, kh, kw), (1, stride[1], stride[2]), dilation,
                                           (0, padding[1], padding[2]), bias, alpha)
            self.temporal_conv = Conv3dBlock(mid, out_channels, (kt, 1, 1), (stride[0], 1, 1), dilation,
                                             (padding[0], 0, 0), bias, alpha)

    def forward(self, x: torch.Tensor):
        x, was_internal = _enter(x)
        x = _call(self.spatio_conv, x)
        x = _call(self.temporal_conv, x)
        return _leave(x, was_internal)


class SpatioTemporalResBlock(nn.Module):
    """lrelu(shortcut(x) + conv2(conv1(x)))   (reference :164-187)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[Tuple[int, int, int], int] = (3, 1, 1),
                 downsample: bool = False, dilation: int = 1, alpha: float = 0.01):
        super().__init__()
        self.downsample = downsample
        padding = kernel_size // 2
        if self.downsample:
            self.downsample_conv = SpatioTemporalConv(in_channels, out_channels, kernel_size=1, stride=2,
                                                      dilation=dilation, padding=0)
            self.conv1 = SpatioTemporalConv(in_channels, out_channels, kernel_size, stride=(2, 2, 2),
                                            dilation=dilation, padding=padding)
        else:
            self.conv1 = SpatioTemporalConv(in_channels, out_channels, kernel_size, stride=(1, 1, 1),
                                            dilation=dilation, padding=padding)
        self.conv2 = SpatioTemporalConv(out_channels, out_channels, kernel_size, stride=(1, 1, 1), padding=padding,
                                        dilation=dilation)
        self.relu = nn.LeakyReLU(alpha)
        self._out_channels = out_channels

    def _dp_layers(self) -> List[Conv3dBlock]:
        ls = [self.conv1.spatio_conv, self.conv1.temporal_conv, self.conv2.spatio_conv, self.conv2.temporal_conv]
        if self.downsample:
            ls += [self.downsample_conv.spatio_conv, self.downsample_conv.temporal_conv]
        return ls

    def forward(self, x: torch.Tensor):
        x, was_internal = _enter(x)
        if _any_inner_hooks(self):
            # module-by-module so hooks registered on inner modules observe their tensors
            res = _call(self.conv2, _call(self.conv1, x))
            sc = _call(self.downsample_conv, x) if self.downsample else x
            out = Fn.AddActFn.apply(res, sc, float(self.relu.negative_slope))
        else:
            layers = self._dp_layers()
            params = []
            for m in layers:
                m._refresh_cfg()
                params += [m.conv.weight, m.bn.weight, m.bn.bias]
            out = Fn.res_block(x, self, *params)
            if self.training and not _COUNTED[0]:
                for m in layers:
                    if m.bn.track_running_stats:
                        m.bn.num_batches_tracked.add_(1)
        Fn.tag(out, self._out_channels)
        return _leave(out, was_internal)


class SpatioTemporalResLayer(nn.Module):
    """block1 (optionally down-sampling) followed by layer_size-1 identity blocks (reference :190-204)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[Tuple[int, int, int], int] = (3, 1, 1),
                 downsample: bool = False, dilation: int = 1, alpha: float = 0.01, layer_size: int = 4):
        super().__init__()
        self.block1 = SpatioTemporalResBlock(in_channels, out_channels, kernel_size, downsample=downsample,
                                             dilation=dilation, alpha=alpha)
        self.blocks = nn.ModuleList([])
        for _ in range(layer_size - 1):
            self.blocks.append(SpatioTemporalResBlock(out_channels, out_channels, kernel_size, downsample=False,
                                                      dilation=dilation, alpha=alpha))

    def forward(self, x: torch.Tensor):
        x, was_internal = _enter(x)
        x = _call(self.block1, x)
        for block in self.blocks:
            x = _call(block, x)
        return _leave(x, was_internal)


class R2Plus1DNet(nn.Module):
    """Stem + four residual stages + global average pool -> (B,128)   (reference :207-226)."""

    def __init__(self, layer_sizes: List[int] = [4, 4, 4, 4], alpha: float = 0.01):
        super().__init__()
        self.conv1 = SpatioTemporalConv(3, 32, kernel_size=(1, 7, 7), stride=(1, 2, 2), padding=(0, 3, 3), dilation=1,
                                        is_first=True, alpha=alpha)
        self.conv2 = SpatioTemporalResLayer(32, 32, 3, dilation=1, alpha=alpha, layer_size=layer_sizes[0])
        self.conv3 = SpatioTemporalResLayer(32, 64, 3, dilation=1, alpha=alpha, layer_size=layer_sizes[1],
                                            downsample=True)
        self.conv4 = SpatioTemporalResLayer(64, 64, 3, dilation=1, alpha=alpha, layer_size=layer_sizes[2],
                                            downsample=True)
        self.conv5 = SpatioTemporalResLayer(64, 128, 3, dilation=1, alpha=alpha, layer_size=layer_sizes[3],
                                            downsample=True)
        self.pool = nn.AdaptiveAvgPool3d(1)
        self.out_features = 128

    def forward(self, x: torch.Tensor, mean_bgr=None):
        """x: (B,3,T,H,W) fp32 clips as the reference's DataLoader yields them, or -- an extension for the
        sliding-window loop -- (B,T,H,W,3) uint8 BGR frames with `mean_bgr` (dataset.py:104-110)."""
        batch_size = x.size(0)
        # nn.BatchNorm3d counts its training batches (num_batches_tracked, part of the reference's state dict): one
        # multi-tensor increment for all layers instead of a 4-us kernel between every conv and the next (32 per step)
        counted = self.training and not _COUNTED[0]
        if counted:
            bns = [m for m in self.modules() if isinstance(m, nn.BatchNorm3d) and m.track_running_stats]
            if bns:
                torch._foreach_add_([m.num_batches_tracked for m in bns], 1)
            _COUNTED[0] = True
        try:
            return self._forward(x, mean_bgr, batch_size)
        finally:
            if counted:
                _COUNTED[0] = False

    def _forward(self, x: torch.Tensor, mean_bgr, batch_size: int):
        stem = self.conv1.spatio_conv
        geom = None if (_hooked(self.conv1) or _hooked(stem) or _hooked(self.conv1.temporal_conv)) else stem._stem_geom(x)
        if geom is not None:
            x = _call(self.conv1.temporal_conv, stem.forward_stem(x, geom, mean_bgr))
            stages = (self.conv2, self.conv3, self.conv4, self.conv5)
        else:
            if x.dtype == torch.uint8:
                x = Fn.frames_u8_to_internal(x, mean_bgr)
            x, _ = _enter(x)
            stages = (self.conv1, self.conv2, self.conv3, self.conv4, self.conv5)
        for stage in stages:
            x = _call(stage, x)
        x = Fn.AvgPoolFn.apply(x, x._dp_c)
        return x.view(batch_size, -1)


class R2Plus1DClassifier(nn.Module):
    """Encoder + MLP head   (reference :228-288)."""

    def __init__(self, input_size: Tuple[int, int, int, int] = (3, 8, 112, 112), num_classes: int = 2,
                 layer_sizes: List[int] = [4, 4, 4, 4], pretrained: bool = False, alpha: float = 1.0):
        super().__init__()
        self.input_size = input_size
        self.res2plus1d = R2Plus1DNet(layer_sizes, alpha=alpha)
        linear_dims = self.get_res2plus1d_output_size()[1]
        self.linear = nn.Sequential(
            nn.Linear(linear_dims, linear_dims // 2),
            nn.BatchNorm1d(linear_dims // 2),
            nn.ELU(alpha),
            nn.Linear(linear_dims // 2, num_classes),
        )
        self._init_weight()
        if pretrained:
            self._load_pretrained_weights()

    def get_res2plus1d_output_size(self):
        # The reference pushes a zero clip through the encoder on the CPU (:255-259); the pooled width
        # does not depend on the clip, so it is stated instead of executed (no CPU path here).
        # That probe runs in training mode, so in the reference every BatchNorm3d leaves the constructor
        # having seen one all-zero batch (bias-free convs of zeros are zeros): running_mean *= (1-momentum),
        # running_var = (1-momentum)*1 + momentum*0 = 0.9, num_batches_tracked = 1.  Reproduced here so a
        # freshly constructed model has the reference's buffers bit for bit.
        self._check_input_size()
        with torch.no_grad():
            for m in self.res2plus1d.modules():
                if isinstance(m, nn.BatchNorm3d) and m.track_running_stats and m.momentum is not None:
                    m.running_mean.mul_(1.0 - m.momentum)
                    m.running_var.mul_(1.0 - m.momentum)
                    m.num_batches_tracked.add_(1)
        return torch.Size([1, self.res2plus1d.out_features])

    def _check_input_size(self):
        c, t, h, w = self.input_size
        if c != 3:
            raise ValueError("R2Plus1DClassifier expects 3-channel clips (C,T,H,W)")
        # every stage must keep at least one pixel: stem /2 spatial, three stages /2 in t,h,w
        th, hh, ww = t, (h + 2 * 3 - 7) // 2 + 1, (w + 2 * 3 - 7) // 2 + 1
        for _ in range(3):
            th, hh, ww = (th - 1) // 2 + 1, (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
        if min(th, hh, ww) < 1:
            raise ValueError(f"input_size {self.input_size} is too small for R(2+1)D")

    def _load_pretrained_weights(self):
        s_dict = self.state_dict()
        for name in s_dict:
            print(name)
            print(s_dict[name].size())

    def _init_weight(self):
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def encode(self, x: torch.Tensor):
        with torch.no_grad():
            x = self.res2plus1d(x)
        return x

    def forward(self, x: torch.Tensor, mean_bgr=None):
        """x: (B,3,T,H,W) fp32 clips (the reference's call), or (B,T,H,W,3) uint8 BGR frames with `mean_bgr`
        (the uint8 input boundary: mean subtraction and layout change run on the device, dataset.py:104-110)."""
        x = self.res2plus1d(x, mean_bgr) if mean_bgr is not None else self.res2plus1d(x)
        x = self.linear(x)
        return x

    def summary(self, device: str = 'cpu', show_input: bool = True, show_hierarchical: bool = True,
                print_summary: bool = False, show_parent_layers: bool = False):
        n_params = sum(p.numel() for p in self.parameters())
        try:
            from pytorch_model_summary import summary as _summary
        except ImportError:
            return print(f"{self}\ninput (B,{','.join(map(str, self.input_size))}) | parameters: {n_params:,}")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            return print(f"{self}\ninput (B,{','.join(map(str, self.input_size))}) | parameters: {n_params:,}")
        sample = torch.zeros((8, *self.input_size), device=dev)
        return print(_summary(self, sample, show_input=show_input, show_hierarchical=show_hierarchical,
                              print_summary=print_summary, show_parent_layers=show_parent_layers))
