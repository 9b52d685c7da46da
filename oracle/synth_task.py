"""ORACLE -- test infrastructure, not product code.

A LEARNABLE synthetic disruption task for convergence A/B runs (tests/test_gpu_convergence.py,
oracle/make_convergence_golden.py): IVIS-like clips (dark vessel, a few bright moving Gaussian blobs, sensor
noise, uint8 BGR minus the dataset mean -- /root/reference/src/dataset.py:104-110,201-205) whose label is carried
by a TEMPORAL feature, as in the real data: in class 0 ("disruption", label 0 as in src/dataset.py:91-94) the
blobs' brightness collapses over the last third of the clip (a thermal-quench-like fade) while their position
jitters; in class 1 ("normal") the brightness stays steady.  Everything is drawn from one torch.Generator seeded
with (seed, step), so the port (CPU, in the build container) and the CUDA path (GPU box) see identical clips.
"""
from __future__ import annotations

import torch

MEAN_BGR = (90.0, 98.0, 102.0)


def task_batch(step: int, B: int, T: int = 21, H: int = 128, W: int = 128, seed: int = 777, noise: int = 6):
    """Returns (x, y): x (B,3,T,H,W) fp32 mean-subtracted uint8-valued clips, y (B,) int64 labels (balanced)."""
    g = torch.Generator().manual_seed(seed * 1000003 + step)
    yy = torch.linspace(0, 1, H).view(1, 1, H, 1)
    xx = torch.linspace(0, 1, W).view(1, 1, 1, W)
    tt = torch.arange(T, dtype=torch.float32).view(1, T, 1, 1)
    y = (torch.arange(B) + int(torch.randint(0, 2, (1,), generator=g))) % 2
    y = y[torch.randperm(B, generator=g)]
    clips = torch.zeros(B, 3, T, H, W)
    for b in range(B):
        img = torch.rand(1, generator=g).item() * 40.0 + 10.0 + torch.zeros(3, T, H, W)
        # class 0: brightness ramps down to 15-45 % over the last third; class 1: stays within +-8 %
        if int(y[b]) == 0:
            t0 = T * (0.55 + 0.15 * torch.rand(1, generator=g).item())
            floor = 0.15 + 0.3 * torch.rand(1, generator=g).item()
            env = 1.0 - (1.0 - floor) * ((tt - t0) / (T - 1 - t0)).clamp(0, 1)
            jit = 0.012
        else:
            env = 1.0 + 0.08 * torch.sin(tt * (0.2 + 0.5 * torch.rand(1, generator=g).item()))
            jit = 0.0
        for _ in range(3):
            cx, cy = (torch.rand(2, generator=g) * 0.6 + 0.2).tolist()
            vx, vy = ((torch.rand(2, generator=g) - 0.5) * 0.02).tolist()
            sig = torch.rand(1, generator=g).item() * 0.15 + 0.06
            amp = torch.rand(1, generator=g).item() * 150.0 + 60.0
            gain = torch.rand(3, generator=g) * 0.4 + 0.6
            jx = jit * torch.randn(T, generator=g).view(1, T, 1, 1) * (tt > T * 0.6)
            r2 = (xx - (cx + vx * tt + jx)) ** 2 + (yy - (cy + vy * tt)) ** 2
            img = img + gain.view(3, 1, 1, 1) * (amp * env * torch.exp(-r2 / (2 * sig * sig)))
        clips[b] = img
    if noise > 0:
        clips += torch.randint(-noise, noise + 1, clips.shape, generator=g).float()
    x = clips.round().clamp_(0, 255)
    x -= torch.tensor(MEAN_BGR).view(1, 3, 1, 1, 1)
    return x, y
