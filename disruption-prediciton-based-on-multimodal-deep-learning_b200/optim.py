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
        self.capturable = capturable   # kept for interface compatibility: the step count ALWAYS lives on the device
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
            "norm": torch.zeros(1, dtype=torch.float32, device=dev),
            "ws": old["ws"] if old.get("n") == n else torch.zeros(int(L.load().dp_optim_workspace(n)), dtype=torch.uint8, device=dev),
            "ptrs": [p.data_ptr() for p in ps],
        }
        self._flat[gi] = st
        return st

    # ------------------------------------------------------------------------------------------------
    # gradients
    # ------------------------------------------------------------------------------------------------
    def grad_buffer(self, gi: int = 0):
        """Make every parameter's `.grad` a view of this optimiser's flat gradient bucket (parameters() order) and
        return (flat, params, offsets).  Autograd then accumulates in place and `step()` needs no gather copy; the
        data-parallel reducer all-reduces contiguous slices of the same buffer (distributed.BucketedGradAllReduce)."""
        st = self._ensure_flat(gi, self.param_groups[gi])
        g = st["grad"]
        for p, o in zip(st["params"], st["offs"]):
            view = g[o:o + p.numel()].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view
        return g, st["params"], st["offs"]

    def flat_grad(self, gi: int = 0):
        """Returns (flat gradient bucket, [(lo, hi)] ranges that received a gradient).  Parameters whose `.grad`
        already is a view of the bucket cost nothing; others are gathered with one fused copy.  Parameters with
        `.grad is None` are left out of the ranges: like torch.optim.AdamW they are neither decayed nor moved."""
        group = self.param_groups[gi]
        st = self._ensure_flat(gi, group)
        g = st["grad"]
        base = g.data_ptr()
        views, srcs, ranges = [], [], []
        for p, o in zip(st["params"], st["offs"]):
            n = p.numel()
            if p.grad is None:
                g[o:o + n].zero_()      # keeps the norm right; the range is skipped below
                continue
            if p.grad.data_ptr() != base + 4 * o:
                views.append(g[o:o + n].view_as(p))
                srcs.append(p.grad)
            hi = o + (n + 3) // 4 * 4
            if ranges and ranges[-1][1] == o:
                ranges[-1] = (ranges[-1][0], hi)
            else:
                ranges.append((o, hi))
        if views:
            torch._foreach_copy_(views, srcs)
        return g, ranges

    # ------------------------------------------------------------------------------------------------
    # step
    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = L.load()
        sp = L.stream_ptr()
        for gi, group in enumerate(self.param_groups):
            st = self._ensure_flat(gi, group)
            g, ranges = self.flat_grad(gi)
            if not ranges:
                continue
            b1, b2 = group["betas"]
            mn = group["max_norm"]
            # the step count lives on the device (OptWs.step): a step whose gradient norm is not finite is skipped
            # there -- weights, moments and the count stay untouched (reference train.py:55-60) -- without a host sync
            L.check(lib.dp_grad_sqnorm(g.data_ptr(), st["n"], float(self.grad_scale), 1, st["norm"].data_ptr(),
                                       st["ws"].data_ptr(), sp), "dp_grad_sqnorm")
            for lo, hi in ranges:
                L.check(lib.dp_adamw_apply(st["flat"].data_ptr() + 4 * lo, g.data_ptr() + 4 * lo,
                                           st["m"].data_ptr() + 4 * lo, st["v"].data_ptr() + 4 * lo, hi - lo,
                                           float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                           float(group["weight_decay"]), 0, float(mn) if mn else 0.0,
                                           float(self.grad_scale), st["norm"].data_ptr(), st["ws"].data_ptr(), sp),
                        "dp_adamw_apply")
            self.last_grad_norm = st["norm"]
        Fn.bump_weight_epoch()   # the kernel wrote the master weights through raw pointers
        return loss

    def _ws_words(self, gi: int = 0) -> torch.Tensor:
        """int32 view of the device workspace header: [counter, step, skipped, skip_now] (csrc/optim.cu OptWs)."""
        return self._flat[gi]["ws"][:16].view(torch.int32)

    def device_step(self, gi: int = 0) -> int:
        """Number of updates applied so far (host sync)."""
        return int(self._ws_words(gi)[1].item()) if gi in self._flat else 0

    def skipped_steps(self, gi: int = 0) -> int:
        """Number of steps skipped because the gradient norm was not finite (host sync)."""
        return int(self._ws_words(gi)[2].item()) if gi in self._flat else 0

    # ------------------------------------------------------------------------------------------------
    # checkpointing: the same layout torch.optim.AdamW writes ({"step", "exp_avg", "exp_avg_sq"} per parameter)
    # ------------------------------------------------------------------------------------------------
    def state_dict(self):
        for gi, group in enumerate(self.param_groups):
            if not any(p.requires_grad for p in group["params"]):
                continue
            st = self._ensure_flat(gi, group)
            step = torch.tensor(float(self.device_step(gi)))
            for p, o in zip(st["params"], st["offs"]):
                n = p.numel()
                self.state[p] = {"step": step.clone(), "exp_avg": st["m"][o:o + n].view_as(p).clone(),
                                 "exp_avg_sq": st["v"][o:o + n].view_as(p).clone()}
        sd = super().state_dict()
        self.state.clear()      # the flat buffers stay the single source of truth
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for gi, group in enumerate(self.param_groups):
            if not any(p.requires_grad for p in group["params"]):
                continue
            st = self._ensure_flat(gi, group)
            step = 0
            for p, o in zip(st["params"], st["offs"]):
                ps = self.state.get(p)
                if not ps:
                    continue
                n = p.numel()
                st["m"][o:o + n].copy_(ps["exp_avg"].reshape(-1))
                st["v"][o:o + n].copy_(ps["exp_avg_sq"].reshape(-1))
                step = max(step, int(float(ps["step"])))
            self._ws_words(gi)[1] = step
        self.state.clear()
