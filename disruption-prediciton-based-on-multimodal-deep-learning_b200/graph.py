"""Whole-step CUDA graph: forward + loss + backward + clip + AdamW captured once and replayed.

A training step of this path is ~370 short kernels (32 conv layers x (conv, BN finalize, BN apply) forward, twice
that backward).  Issued from Python they leave the GPU idle ~15 % of the step; replayed as one CUDA graph the host
cost is a single launch.  Opt-in: the reference's `train_per_epoch` (/root/reference/src/train.py:38-75) reads the
loss on the host three times per step and cannot be captured, so this is the step body of that loop
(zero_grad -> model -> loss -> backward -> clip -> optimizer.step) as one replayable object.

Requirements: static shapes, `FusedClipAdamW(capturable=True)` (step count on the device), no host
synchronisation inside the step.  The packed bf16 weight copies are rebuilt inside the graph every step (their
pack kernels are captured), so replay sees the weights the optimiser just wrote.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib as L
from . import functional as Fn
from .optim import FusedClipAdamW


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer: FusedClipAdamW, example_x: torch.Tensor,
                 example_y: torch.Tensor, warmup: int = 3, pre_backward: Optional[Callable] = None,
                 post_backward: Optional[Callable] = None, forward: Optional[Callable] = None):
        """`pre_backward` / `post_backward` run inside the captured step around `loss.backward()` (the data-parallel
        trainer zeroes its gradient buckets / waits for the all-reduces there).  `forward` replaces `model(x)` when the
        batch is not what `model.forward` takes (uint8 frames + BGR mean: `lambda x: model(x, mean_bgr=MEAN)`)."""
        L.require_device()
        if not getattr(optimizer, "capturable", False):
            raise L.DpError("GraphedTrainStep needs FusedClipAdamW(capturable=True)")
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.forward = forward if forward is not None else model
        # several input tensors (video clip + 0D signals of the multimodal model) are passed as a tuple
        self._multi = isinstance(example_x, (tuple, list))
        self.static_x = tuple(t.clone() for t in example_x) if self._multi else example_x.clone()
        self.static_y = example_y.clone()
        self.pre_backward, self.post_backward = pre_backward, post_backward
        # AccumulateGrad nodes created by earlier eager steps live on the default stream for as long as any tensor
        # (e.g. an old loss) keeps their autograd graph alive, and would drag the legacy stream into the capture:
        # the caller must drop such references; collect what is already unreachable
        import gc
        gc.collect()
        side = torch.cuda.Stream(device=example_y.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # allocator, workspaces, weight packs, kernel attributes settle
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        lib = L.load()
        n0 = lib.dp_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss, self.static_out = self._body()
        self.launches_per_step = int(lib.dp_launch_count() - n0)   # kernels of this library inside one replay

    def _body(self):
        if self.pre_backward is not None:
            self.pre_backward()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        out = self.forward(*self.static_x) if self._multi else self.forward(self.static_x)
        # a tuple of heads (the three-head GradientBlending model, train.py:46-50) goes to the loss unpacked
        loss = self.loss_fn(*out, self.static_y) if isinstance(out, tuple) else self.loss_fn(out, self.static_y)
        loss.backward()
        if self.post_backward is not None:
            self.post_backward()
        self.optimizer.step()
        return loss.detach(), (tuple(o.detach() for o in out) if isinstance(out, tuple) else out.detach())

    def step(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None):
        """Copy the batch into the static buffers (any device/pinned-host source, async) and replay.
        Returns (loss, logits) as static device tensors valid until the next step."""
        if x is not None and self._multi:
            for dst, src in zip(self.static_x, x):
                dst.copy_(src, non_blocking=True)
        elif x is not None:
            self.static_x.copy_(x, non_blocking=True)
        if y is not None:
            self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        Fn.bump_weight_epoch()   # the replay rewrote the master weights: eager callers must re-pack
        return self.static_loss, self.static_out
