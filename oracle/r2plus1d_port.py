"""ORACLE -- test infrastructure, not product code.

CPU restatement (PyTorch fp32 ops, functional style, no nn.Module tree) of the reference's hot path:
the R(2+1)D encoder + head (/root/reference/src/models/R2Plus1D.py:25-288), the three losses
(/root/reference/src/loss.py:14-81) and RW / DRW class weights (train_vision_network.py:312-318,
src/train.py:318-329).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it; the product package never does.

Where the arithmetic lives: the reference calls third-party PyTorch (torch.nn.functional.conv3d,
batch_norm, leaky_relu, cross_entropy; no version pinned by the reference -- README cites a missing
environment.yaml; this container and the GPU box run torch 2.11.0).  This port calls the same
primitives in the same order on a plain {name: tensor} state dict with the reference's key names, so
the same seeded state gives the same numbers.

Pinning: the reference's own tests hold no golden vectors for this path, so the
port is pinned against the reference itself, imported in the build container by
oracle/make_golden.py, through the fixtures in tests/golden/ (tests/test_oracle_golden.py).  An
independent numpy restatement of the primitives is in oracle/np_ops.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
INNER_SLOPE = 0.01  # SpatioTemporalConv default alpha, used inside res blocks (R2Plus1D.py:116,172-177)


# ----------------------------------------------------------------------------------------------
# layer plan
# ----------------------------------------------------------------------------------------------
def mid_channels(cin: int, cout: int, k: Tuple[int, int, int]) -> int:
    """R2Plus1D.py:150-155."""
    kt, kh, kw = k
    return int(math.floor((kt * kh * kw * cin * cout) / (kh * kw * cin + kt * cout)))


def st_conv_plan(prefix: str, cin: int, cout: int, k, stride, padding, slope: float, is_first: bool = False):
    """Two Conv3dBlocks of one SpatioTemporalConv: (name, cin, cout, kernel, stride, padding, slope)."""
    kt, kh, kw = k
    if is_first:
        mid = 45
        return [
            (prefix + ".spatio_conv", cin, mid, (kt, kh, kw), (1, stride[1], stride[2]), tuple(padding), slope),
            (prefix + ".temporal_conv", mid, cout, (3, 1, 1), (stride[0], 1, 1), (1, 0, 0), slope),
        ]
    mid = mid_channels(cin, cout, k)
    return [
        (prefix + ".spatio_conv", cin, mid, (1, kh, kw), (1, stride[1], stride[2]), (0, padding[1], padding[2]), slope),
        (prefix + ".temporal_conv", mid, cout, (kt, 1, 1), (stride[0], 1, 1), (padding[0], 0, 0), slope),
    ]


def encoder_plan(layer_sizes: Sequence[int], alpha: float):
    """Returns (stem_layers, blocks); block = dict(prefix, conv1, conv2, shortcut|None, slope)."""
    stem = st_conv_plan("res2plus1d.conv1", 3, 32, (1, 7, 7), (1, 2, 2), (0, 3, 3), alpha, is_first=True)
    blocks = []
    chans = [(32, 32, False), (32, 64, True), (64, 64, True), (64, 128, True)]
    for si, ((cin, cout, ds), n) in enumerate(zip(chans, layer_sizes)):
        stage = f"res2plus1d.conv{si + 2}"
        for bi in range(n):
            prefix = f"{stage}.block1" if bi == 0 else f"{stage}.blocks.{bi - 1}"
            bin_, down = (cin, ds) if bi == 0 else (cout, False)
            s = (2, 2, 2) if down else (1, 1, 1)
            blocks.append(dict(
                prefix=prefix,
                conv1=st_conv_plan(prefix + ".conv1", bin_, cout, (3, 3, 3), s, (1, 1, 1), INNER_SLOPE),
                conv2=st_conv_plan(prefix + ".conv2", cout, cout, (3, 3, 3), (1, 1, 1), (1, 1, 1), INNER_SLOPE),
                shortcut=st_conv_plan(prefix + ".downsample_conv", bin_, cout, (1, 1, 1), (2, 2, 2), (0, 0, 0),
                                      INNER_SLOPE) if down else None,
                slope=alpha,
            ))
    return stem, blocks


def all_conv_layers(layer_sizes: Sequence[int], alpha: float):
    stem, blocks = encoder_plan(layer_sizes, alpha)
    out = list(stem)
    for b in blocks:
        out += b["conv1"] + b["conv2"] + (b["shortcut"] or [])
    return out


# ----------------------------------------------------------------------------------------------
# state
# ----------------------------------------------------------------------------------------------
def init_state(layer_sizes: Sequence[int], num_classes: int = 2, seed: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """A state dict with the reference's key names and its final initialisation DISTRIBUTIONS
    (kaiming-normal conv weights, BN gamma=1/beta=0, default Linear init).  It does not try to
    reproduce the reference's RNG stream; tests that need identical weights pass a state dict."""
    g = torch.Generator().manual_seed(0 if seed is None else seed)
    st: Dict[str, torch.Tensor] = {}
    for name, cin, cout, k, _, _, _ in all_conv_layers(layer_sizes, 1.0):
        fan_in = cin * k[0] * k[1] * k[2]
        st[name + ".conv.weight"] = torch.randn((cout, cin, *k), generator=g) * math.sqrt(2.0 / fan_in)
        st[name + ".bn.weight"] = torch.ones(cout)
        st[name + ".bn.bias"] = torch.zeros(cout)
        st[name + ".bn.running_mean"] = torch.zeros(cout)
        st[name + ".bn.running_var"] = torch.ones(cout)
        st[name + ".bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    for i, (fi, fo) in ((0, (128, 64)), (3, (64, num_classes))):
        bound = 1.0 / math.sqrt(fi)
        st[f"linear.{i}.weight"] = (torch.rand((fo, fi), generator=g) * 2 - 1) * bound
        st[f"linear.{i}.bias"] = (torch.rand((fo,), generator=g) * 2 - 1) * bound
    st["linear.1.weight"] = torch.ones(64)
    st["linear.1.bias"] = torch.zeros(64)
    st["linear.1.running_mean"] = torch.zeros(64)
    st["linear.1.running_var"] = torch.ones(64)
    st["linear.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return st


def clone_state(state: Dict[str, torch.Tensor], requires_grad: bool = True, dtype=None) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in state.items():
        t = v.detach().clone().cpu()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


# ----------------------------------------------------------------------------------------------
# bf16-storage emulation
# ----------------------------------------------------------------------------------------------
# storage="bf16" restates the SAME fp32 algorithm but rounds tensors to bfloat16 at exactly the
# points where the CUDA product path stores them in HBM (conv input, packed weights, raw conv output,
# post-activation output, and the gradients of those tensors on the way back).  It answers "what does
# the reference's algorithm give when activations live in bf16", which is the fair checker for the bf16
# kernels: bf16 storage alone moves logits by several percent on noise inputs (numerics).
class _RoundBoth(torch.autograd.Function):
    """Round to bf16 in forward (stored activation) and in backward (stored gradient)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().to(g.dtype)


class _RoundFwd(torch.autograd.Function):
    """Round to bf16 in forward only (packed weights; weight gradients stay fp32)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _q(x, storage):
    return _RoundBoth.apply(x) if storage == "bf16" else x


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def conv_bn_act(st, name, x, kernel, stride, padding, slope, training, taps=None, storage="fp32", round_out=True):
    """Conv3dBlock.forward: LeakyReLU(BN3d(Conv3d(x)))   (R2Plus1D.py:56-58)."""
    w = st[name + ".conv.weight"]
    if storage == "bf16":
        w = _RoundFwd.apply(w)
    y = _q(F.conv3d(x, w, None, stride, padding), storage)
    rm, rv = st[name + ".bn.running_mean"], st[name + ".bn.running_var"]
    z = F.batch_norm(y, rm, rv, st[name + ".bn.weight"], st[name + ".bn.bias"], training, BN_MOMENTUM, BN_EPS)
    if training:
        st[name + ".bn.num_batches_tracked"] += 1
    out = F.leaky_relu(z, slope)
    if round_out:
        out = _q(out, storage)
    if taps is not None:
        taps[name + ".conv"] = y
        taps[name] = out
    return out


def st_conv(st, layers, x, training, taps=None, storage="fp32", round_last=True):
    for i, (name, _, _, k, s, p, slope) in enumerate(layers):
        x = conv_bn_act(st, name, x, k, s, p, slope, training, taps, storage,
                        round_out=round_last or i + 1 < len(layers))
    return x


def encoder_forward(st, x, layer_sizes, alpha, training=True, taps=None, storage="fp32"):
    """R2Plus1DNet.forward (R2Plus1D.py:217-226): x (B,3,T,H,W) fp32 -> (B,128)."""
    stem, blocks = encoder_plan(layer_sizes, alpha)
    x = _q(x, storage)
    x = st_conv(st, stem, x, training, taps, storage)
    for b in blocks:
        res = st_conv(st, b["conv1"], x, training, taps, storage)
        # the CUDA path keeps conv2's activation in registers and stores only lrelu(shortcut + res)
        res = st_conv(st, b["conv2"], res, training, taps, storage, round_last=False)
        if b["shortcut"] is not None:
            sc = st_conv(st, b["shortcut"], _q(x, storage), training, taps, storage)
        else:
            sc = _q(x, storage)
        x = _q(F.leaky_relu(sc + res, b["slope"]), storage)
        if taps is not None:
            taps[b["prefix"]] = x
    x = F.adaptive_avg_pool3d(x, 1)
    return x.view(x.size(0), -1)


def head_forward(st, feat, alpha, training=True):
    """Linear -> BatchNorm1d -> ELU(alpha) -> Linear   (R2Plus1D.py:243-248)."""
    h = F.linear(feat, st["linear.0.weight"], st["linear.0.bias"])
    h = F.batch_norm(h, st["linear.1.running_mean"], st["linear.1.running_var"], st["linear.1.weight"],
                     st["linear.1.bias"], training, BN_MOMENTUM, BN_EPS)
    if training:
        st["linear.1.num_batches_tracked"] += 1
    h = F.elu(h, alpha)
    return F.linear(h, st["linear.3.weight"], st["linear.3.bias"])


def classifier_forward(st, x, layer_sizes, alpha, training=True, taps=None, storage="fp32"):
    return head_forward(st, encoder_forward(st, x, layer_sizes, alpha, training, taps, storage), alpha, training)


# ----------------------------------------------------------------------------------------------
# losses and class weights
# ----------------------------------------------------------------------------------------------
def ce_rows(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return torch.logsumexp(logits, dim=1) - logits.gather(1, target.view(-1, 1)).squeeze(1)


def focal_loss(logits, target, weight, gamma: float = 2.0):
    """loss.py:25-34: sum_i w[y_i] (1 - exp(-CE_i))^gamma CE_i."""
    ce = ce_rows(logits, target)
    p = torch.exp(-ce)
    return (weight.to(logits.dtype)[target] * (1 - p) ** gamma * ce).sum()


def ldam_margins(cls_num_list, max_m: float = 0.5) -> torch.Tensor:
    """loss.py:52-56."""
    m = 1.0 / np.sqrt(np.sqrt(np.asarray(cls_num_list, dtype=np.float64)))
    m = m * (max_m / np.max(m))
    return torch.tensor(m, dtype=torch.float32)


def ldam_loss(logits, target, margins, weight=None, s: float = 30.0):
    """loss.py:58-69: weighted-mean CE of s * (z - m[y] onehot(y))."""
    onehot = F.one_hot(target, logits.shape[1]).to(logits.dtype)
    z = s * (logits - onehot * margins.to(logits.dtype)[target].view(-1, 1))
    ce = ce_rows(z, target)
    if weight is None:
        return ce.mean()
    w = weight.to(logits.dtype)[target]
    return (w * ce).sum() / w.sum()


def ce_loss(logits, target, weight=None):
    """loss.py:80-81: weighted CE, reduction='sum'."""
    ce = ce_rows(logits, target)
    if weight is None:
        return ce.sum()
    return (weight.to(logits.dtype)[target] * ce).sum()


def rw_class_weights(cls_num_list, use_weighting=True) -> torch.Tensor:
    """train_vision_network.py:312-318."""
    if not use_weighting:
        return torch.tensor([1.0, 1.0])
    w = 1.0 / np.asarray(cls_num_list, dtype=np.float64)
    return torch.tensor(w / w.sum(), dtype=torch.float32)


def drw_class_weights(epoch, num_epoch, betas, cls_num_list) -> torch.Tensor:
    """src/train.py:318-329."""
    idx = min(epoch // int(num_epoch / len(betas)), len(betas) - 1)
    beta = betas[idx]
    eff = 1.0 - np.power(beta, np.asarray(cls_num_list, dtype=np.float64))
    w = (1.0 - beta) / eff
    w = w / w.sum() * len(cls_num_list)
    return torch.tensor(w, dtype=torch.float32)


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synthetic_clips(B: int, T: int = 21, H: int = 128, W: int = 128, seed: int = 1234):
    """uint8-valued frames minus the BGR mean [90,98,102], NCDHW fp32 (src/dataset.py:104-110,201-205,229-230)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (B, 3, T, H, W), generator=g).float()
    x -= torch.tensor([90.0, 98.0, 102.0]).view(1, 3, 1, 1, 1)
    y = torch.randint(0, 2, (B,), generator=g)
    return x, y


def structured_clips(B: int, T: int = 21, H: int = 128, W: int = 128, seed: int = 1234, noise: int = 6):
    """Camera-like synthetic clips: a dark background with a few bright, slowly moving Gaussian blobs whose
    position / size / brightness / colour differ from clip to clip, plus +-`noise` grey levels of sensor noise,
    quantised to uint8 and mean-subtracted exactly like `synthetic_clips`.  IVIS frames are of this kind (a
    bright plasma boundary on a dark vessel); unlike i.i.d. pixel noise, the clips of a batch differ
    macroscopically, so the batch statistics of the head's BatchNorm1d are well conditioned."""
    g = torch.Generator().manual_seed(seed)
    yy = torch.linspace(0, 1, H).view(1, 1, H, 1)
    xx = torch.linspace(0, 1, W).view(1, 1, 1, W)
    tt = torch.arange(T, dtype=torch.float32).view(1, T, 1, 1)
    clips = torch.zeros(B, 3, T, H, W)
    for b in range(B):
        img = torch.rand(1, generator=g).item() * 55.0 + 5.0 + torch.zeros(3, T, H, W)
        for _ in range(4):
            cx, cy = (torch.rand(2, generator=g) * 0.6 + 0.2).tolist()
            vx, vy = ((torch.rand(2, generator=g) - 0.5) * 0.02).tolist()
            sig = torch.rand(1, generator=g).item() * 0.2 + 0.05
            amp = torch.rand(1, generator=g).item() * 180.0 + 40.0
            gain = torch.rand(3, generator=g) * 0.4 + 0.6
            flick = 1.0 + 0.1 * torch.sin(tt * (torch.rand(1, generator=g).item() * 0.8 + 0.1))
            r2 = (xx - (cx + vx * tt)) ** 2 + (yy - (cy + vy * tt)) ** 2
            blob = amp * flick * torch.exp(-r2 / (2 * sig * sig))          # (1,T,H,W)
            img = img + gain.view(3, 1, 1, 1) * blob
        clips[b] = img
    if noise > 0:
        clips += torch.randint(-noise, noise + 1, clips.shape, generator=g).float()
    x = clips.round().clamp_(0, 255)
    x -= torch.tensor([90.0, 98.0, 102.0]).view(1, 3, 1, 1, 1)
    y = torch.randint(0, 2, (B,), generator=g)
    return x, y


def train_step(st, x, y, layer_sizes, alpha, loss="focal", weight=None, margins=None, gamma=2.0, s=30.0,
               storage="fp32"):
    """One forward + loss + backward on the port; returns (logits, loss, {name: grad})."""
    logits = classifier_forward(st, x, layer_sizes, alpha, training=True, storage=storage)
    if weight is None:
        weight = torch.ones(logits.shape[1])
    if loss == "focal":
        l = focal_loss(logits, y, weight, gamma)
    elif loss == "ldam":
        l = ldam_loss(logits, y, margins, weight, s)
    else:
        l = ce_loss(logits, y, weight)
    l.backward()
    grads = {k: v.grad for k, v in st.items() if v.requires_grad and v.grad is not None}
    return logits.detach(), l.detach(), grads
