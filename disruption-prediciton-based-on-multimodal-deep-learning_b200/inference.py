"""Continuous sliding-window disruption prediction (config 5) on the dp_b200 kernels.

Mirrors the loop of `generate_prob_curve` (/root/reference/src/utils/utility.py:896-977) and its twin
`generate_real_time_experiment` (src/visualization/visualize_application.py:190-262): window i of a
shot covers frames i+1 .. i+seq_len (utility.py:404-408), there are len(frames)-seq_len-dist windows
(:402), the model runs in eval mode and the curve is softmax(logits)[:, 0] (label 0 = disruption,
src/dataset.py:91-94).  The reference runs batch 1 and re-reads 21 JPEGs per window; here the shot's
uint8 frames live on the device once, windows are gathered there and batched, and window index
ranges shard across ranks with no collective until the final gather.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import functional as Fn

MEAN_BGR = (90.0, 98.0, 102.0)  # reference src/dataset.py:104-110


def num_windows(n_frames: int, seq_len: int, dist: int) -> int:
    return max(0, n_frames - seq_len - dist)


@torch.no_grad()
def sliding_window_probs(model: torch.nn.Module, frames_u8: torch.Tensor, seq_len: int = 21, dist: int = 3,
                         batch_size: int = 64, window_range: Optional[range] = None,
                         mean_bgr: Sequence[float] = MEAN_BGR) -> torch.Tensor:
    """frames_u8: (N,H,W,3) uint8 BGR frames of one shot on the GPU.  Returns P(disruption) per window."""
    n = num_windows(frames_u8.shape[0], seq_len, dist)
    rng = window_range if window_range is not None else range(n)
    was_training = model.training
    model.eval()
    probs = []
    enc = model.res2plus1d
    try:
        idx0 = torch.arange(1, seq_len + 1, device=frames_u8.device)
        for s in range(rng.start, rng.stop, batch_size):
            e = min(rng.stop, s + batch_size)
            starts = torch.arange(s, e, device=frames_u8.device)
            clip = frames_u8[(starts[:, None] + idx0[None, :])]          # (b, T, H, W, 3) uint8 gather
            feat = enc(clip, mean_bgr)                                   # fused mean-subtract + layout in the stem
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
