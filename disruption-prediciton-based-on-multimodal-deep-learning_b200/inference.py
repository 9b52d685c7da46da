"""Continuous sliding-window disruption prediction (config 5) on the dp_b200 kernels.

Mirrors the loop of `generate_prob_curve` (/root/reference/src/utils/utility.py:896-977) and its twin
`generate_real_time_experiment` (src/visualization/visualize_application.py:190-262): window i of a
shot covers frames i+1 .. i+seq_len (utility.py:404-408), there are len(frames)-seq_len-dist windows
(:402), the model runs in eval mode and the curve is softmax(logits)[:, 0] (label 0 = disruption,
src/dataset.py:91-94).  The reference runs batch 1 and re-reads 21 JPEGs per window; here the shot's
uint8 frames live on the device once; the per-frame stem conv is computed once per FRAME and cached,
windows are an overlapping view of that cache (no gather copy), every layer is one fused kernel
(BatchNorm folded into the conv epilogue), and window index ranges shard across ranks with no
collective until the final gather.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import functional as Fn

MEAN_BGR = (90.0, 98.0, 102.0)  # reference src/dataset.py:104-110


def num_windows(n_frames: int, seq_len: int, dist: int) -> int:
    return max(0, n_frames - seq_len - dist)


def _stem_frame_cache(enc, frames_u8: torch.Tensor, mean_bgr):
    """Stem spatial Conv3dBlock (1x7x7, per frame) over EVERY frame of the shot once: consecutive windows share
    seq_len-1 of their seq_len frames, so the per-window loop of the reference recomputes it ~seq_len times.
    Returns the internal (1,N,Ho,Wo,Cp) activation, or None when the fast path does not apply (fp32 validation mode,
    hooks on the stem, unsupported geometry)."""
    from .R2Plus1D import _hooked
    stem = enc.conv1.spatio_conv
    if _hooked(enc.conv1) or _hooked(stem) or _hooked(enc.conv1.temporal_conv) or Fn.get_compute_mode() != "bf16":
        return None
    shot = frames_u8.unsqueeze(0)                       # (1,N,H,W,3): the shot as one long clip
    geom = stem._stem_geom(shot)
    if geom is None:
        return None
    return stem.forward_stem(shot, geom, mean_bgr)


def _windows_from_cache(enc, cache: torch.Tensor, first: int, count: int, seq_len: int) -> torch.Tensor:
    """Temporal stem conv + the four residual stages + pool for windows first .. first+count-1, each covering cached
    frames i+1 .. i+seq_len (utility.py:404-408).  The windows are an OVERLAPPING VIEW of the frame cache (batch stride
    = time stride = one frame): no gather copy; the view's out-of-range time steps read as zero, which is exactly the
    temporal conv's per-window zero padding."""
    from .R2Plus1D import _call
    m = enc.conv1.temporal_conv
    m._refresh_cfg()
    _, n, ho, wo, cp = cache.shape
    frame = ho * wo * cp
    cfg = m._cfg
    key = ("win", cfg.C, cfg.K, count, seq_len, ho, wo)
    geom = Fn._GEOMS.get(key)
    if geom is None:
        geom = Fn.ConvGeom(cfg.C, cfg.K, cfg.kernel, cfg.stride, cfg.padding, count, seq_len, ho, wo, Fn.L.DP_BF16)
        Fn._GEOMS[key] = geom
    if first + 1 + count - 1 + seq_len > n:
        raise Fn.L.DpError("sliding window runs past the end of the shot")
    z, _ = Fn.layer_forward(cache, m.conv.weight, m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var, cfg,
                            False, cache=m._packed, geom=geom, fuse_eval=True, x_view=(cp, wo * cp, frame, frame),
                            x_ptr=cache.data_ptr() + 2 * frame * (first + 1))
    x = Fn.tag(z, cfg.K)
    for stage in (enc.conv2, enc.conv3, enc.conv4, enc.conv5):
        x = _call(stage, x)
    x = Fn.AvgPoolFn.apply(x, x._dp_c)
    return x.view(count, -1)


class _GraphedWindows:
    """One CUDA graph per (window count, geometry): temporal stem conv -> residual stages -> pool -> head -> softmax over
    a STATIC slab of cached stem frames.  A batch then costs one slab copy, one replay and one read-out instead of ~65
    Python-issued launches -- at batch 1 the reference loop's regime (utility.py:936-949) is launch-bound, not
    compute-bound.  The graph bakes in the packed bf16 weight copies, so it is keyed by the weights' version."""

    def __init__(self, model, cache: torch.Tensor, count: int, seq_len: int):
        enc = model.res2plus1d
        _, _, ho, wo, cp = cache.shape
        self.n_frames = count + seq_len - 1
        self.slab = torch.zeros((1, self.n_frames + 1, ho, wo, cp), dtype=cache.dtype, device=cache.device)
        self.count, self.seq_len = count, seq_len

        def body():
            # slab frame j holds cached frame first + j; _windows_from_cache reads frames (first_arg + 1) + ..., so first_arg = -1
            feat = _windows_from_cache(enc, self.slab, -1, count, seq_len)
            return torch.softmax(model.linear(feat), dim=1)[:, 0]

        self.slab[0, :self.n_frames].copy_(cache[0, :self.n_frames])
        body()                                   # warm-up: weight packs, workspaces, allocator
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = body()

    def run(self, cache: torch.Tensor, first: int) -> torch.Tensor:
        self.slab[0, :self.n_frames].copy_(cache[0, first + 1:first + 1 + self.n_frames])
        self.graph.replay()
        return self.out.clone()


def _weights_key(model) -> tuple:
    return (Fn._WEIGHT_EPOCH[0], Fn.get_compute_mode(), Fn._STATE["impl"],
            sum(p._version for p in model.parameters()), sum(b._version for b in model.buffers()))


@torch.no_grad()
def sliding_window_probs(model: torch.nn.Module, frames_u8: torch.Tensor, seq_len: int = 21, dist: int = 3,
                         batch_size: int = 64, window_range: Optional[range] = None,
                         mean_bgr: Sequence[float] = MEAN_BGR, stem_cache: bool = True, use_graph: bool = True) -> torch.Tensor:
    """frames_u8: (N,H,W,3) uint8 BGR frames of one shot on the GPU.  Returns P(disruption) per window.

    Eval mode runs one fused kernel per layer (conv + BatchNorm running statistics + LeakyReLU [+ residual] in the
    epilogue, functional.layer_forward(fuse_eval=True)); with `stem_cache` the per-frame stem conv runs once per frame
    of the shot instead of once per window and frame."""
    n = num_windows(frames_u8.shape[0], seq_len, dist)
    rng = window_range if window_range is not None else range(n)
    was_training = model.training
    model.eval()
    probs = []
    enc = model.res2plus1d
    try:
        cache = _stem_frame_cache(enc, frames_u8, mean_bgr) if (stem_cache and len(rng) > 0) else None
        idx0 = torch.arange(1, seq_len + 1, device=frames_u8.device)
        graphs = None
        if cache is not None and use_graph:
            key = _weights_key(model)
            store = getattr(model, "_dp_window_graphs", None)
            if store is None or store.get("key") != key:      # weights changed: the captured packed copies are stale
                store = {"key": key}
                model._dp_window_graphs = store
            graphs = store
        for s in range(rng.start, rng.stop, batch_size):
            e = min(rng.stop, s + batch_size)
            if graphs is not None:
                gk = (e - s, seq_len, tuple(cache.shape[2:]))
                g = graphs.get(gk)
                if g is None:
                    g = graphs[gk] = _GraphedWindows(model, cache, e - s, seq_len)
                probs.append(g.run(cache, s))
                continue
            if cache is not None:
                feat = _windows_from_cache(enc, cache, s, e - s, seq_len)
            else:
                starts = torch.arange(s, e, device=frames_u8.device)
                clip = frames_u8[(starts[:, None] + idx0[None, :])]      # (b, T, H, W, 3) uint8 gather
                feat = enc(clip, mean_bgr)                               # fused mean-subtract + layout in the stem
            logits = model.linear(feat)
            probs.append(torch.softmax(logits, dim=1)[:, 0])
    finally:
        model.train(was_training)
    return torch.cat(probs) if probs else torch.empty(0, device=frames_u8.device)


def postprocess_curve(prob_list: List[float], clip_len: int, frame_srt: int, fps: int = 210) -> List[float]:
    """Start-up padding and peaking suppression exactly as utility.py:951-960."""
    out = [0] * (clip_len + frame_srt) + list(prob_list[1:-1])
    for i, p in enumerate(out):
        if i < fps * 1 and p >= 0.5:
            out[i] = 0
    return out
