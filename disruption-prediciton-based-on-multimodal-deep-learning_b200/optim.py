"""Fused optimiser tail: global-norm clip + AdamW over ONE flat fp32 bucket (csrc/optim.cu).

Replaces the per-tensor `clip_grad_norm_` + `AdamW.step` of the reference train loop
(/root/reference/src/train.py:63-66; optimiser built at train_vision_network.py:278 with
`torch.optim.AdamW(model.parameters(), lr=...)`) by two kernel launches and no host synchronisation.
Opt-in: pass it where the reference passes `torch.optim.AdamW`, and give `max_norm` here instead of
`max_norm_grad` to the train loop.  Parameters are re-pointed to views of one flat buffer, so
`state_dict()`, `model.to()`, checkpoints and `zero_grad()` keep working.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _lib as L
from . import functional as Fn


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_norm: Optional[float] = None, capturable: bool = False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm)
        super().__init__(params, defaults)
        self.capturable = capturable   # step count kept on the device: the step can live inside a CUDA graph
        self._flat = {}
        self.grad_scale = 1.0          # set to 1/world_size by the DP trainer (DDP-mean semantics)
        self.last_grad_norm = None     # device scalar tensor, no sync

    def _ensure_flat(self, gi: int, group):
        st = self._flat.get(gi)
        ps = [p for p in group["params"] if p.requires_grad]
        if st is not None and st["ptrs"] == [p.data_ptr() for p in ps]:
            return st
        L.require_device()
        dev = ps[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 for p in ps):
            raise L.DpError("FusedClipAdamW needs CUDA float32 parameters")
        # 4-element alignment per tensor keeps every view 16-byte aligned
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        flat = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, o in zip(ps, offs):
            flat[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[o:o + p.numel()].view_as(p.data)
        old = st or {}
        st = {
            "params": ps, "offs": offs, "n": n, "flat": flat,
            "grad": torch.zeros(n, dtype=torch.float32, device=dev),
            "m": old.get("m") if old.get("n") == n else torch.zeros(n, dtype=torch.float32, device=dev),
            "v": old.get("v") if old.get("n") == n else torch.zeros(n, dtype=torch.float32, device=dev),
            "step": old.get("step", 0) if old.get("n") == n else 0,
            "norm": torch.zeros(1, dtype=torch.float32, device=dev),
            "ws": torch.zeros(int(L.load().dp_optim_workspace(n)), dtype=torch.uint8, device=dev),
            "ptrs": [p.data_ptr() for p in ps],
        }
        self._flat[gi] = st
        return st

    def flat_grad(self, gi: int = 0) -> torch.Tensor:
        """Gather the per-parameter .grad tensors into the flat gradient bucket (one fused copy)."""
        group = self.param_groups[gi]
        st = self._ensure_flat(gi, group)
        g = st["grad"]
        views, srcs = [], []
        for p, o in zip(st["params"], st["offs"]):
            if p.grad is None:
                g[o:o + p.numel()].zero_()
            else:
                views.append(g[o:o + p.numel()].view_as(p))
                srcs.append(p.grad)
        if views:
            torch._foreach_copy_(views, srcs)
        return g

    @torch.no_grad()
    def step(self, closure=None, flat_grad: Optional[torch.Tensor] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        for gi, group in enumerate(self.param_groups):
            st = self._ensure_flat(gi, group)
            g = flat_grad if (flat_grad is not None and gi == 0) else self.flat_grad(gi)
            st["step"] += 1
            b1, b2 = group["betas"]
            mn = group["max_norm"]
            L.check(lib.dp_clip_adamw_step(st["flat"].data_ptr(), g.data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(),
                                           st["n"], float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                           float(group["weight_decay"]), 0 if self.capturable else int(st["step"]),
                                           float(mn) if mn else 0.0, float(self.grad_scale), st["norm"].data_ptr(),
                                           st["ws"].data_ptr(), L.stream_ptr()), "dp_clip_adamw_step")
            self.last_grad_norm = st["norm"]
        Fn.bump_weight_epoch()   # the kernel wrote the master weights through raw pointers
        return loss
