"""Batch-axis data parallelism: one process per GPU, replicated weights, local BatchNorm statistics,
ONE gradient all-reduce (mean) per step over NCCL, launched bucket by bucket while backward is still
running.

The reference's src/distributed.py (mp.spawn -> init_process_group("nccl") -> DDP, :29-129) never
reduces a gradient (it calls the unwrapped model, :74); this module implements the semantics that
file intends: DDP-style MEAN of per-rank gradients, no SyncBatchNorm, identical optimiser step on
every rank.  Launch with torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR from the env).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> tuple:
    """Returns (rank, local_rank, world_size); initialises the process group when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous shard of `n_items` independent units (clips / sliding windows) for this rank."""
    per = (n_items + world - 1) // world
    return range(min(n_items, rank * per), min(n_items, (rank + 1) * per))


def stage_buckets(model: torch.nn.Module) -> List[List[torch.nn.Parameter]]:
    """Parameters grouped in the order backward produces their gradients: head, conv5, ..., conv1."""
    enc = getattr(model, "res2plus1d", None)
    if enc is None:
        return [[p for p in model.parameters() if p.requires_grad]]
    buckets = []
    head = [p for n, p in model.named_parameters() if not n.startswith("res2plus1d.") and p.requires_grad]
    if head:
        buckets.append(head)
    for name in ("conv5", "conv4", "conv3", "conv2", "conv1"):
        ps = [p for p in getattr(enc, name).parameters() if p.requires_grad]
        if ps:
            buckets.append(ps)
    return buckets


class BucketedGradAllReduce:
    """Flat fp32 gradient buckets, all-reduced asynchronously as soon as a bucket is complete.

    `p.grad` of every parameter is a view into its bucket, so autograd writes gradients in place and
    the collective needs no gather copy; `finish()` waits for the collectives and leaves SUMMED
    gradients in place (the optimiser applies 1/world through `grad_scale`, or call with average=True).

    With `optimizer=FusedClipAdamW(...)` the buckets are contiguous slices of the optimiser's own flat
    gradient buffer (parameters() order), so `optimizer.step()` consumes the reduced buckets directly:
    there is no second flat buffer and no copy between the collective and the update.
    """

    def __init__(self, model: torch.nn.Module, buckets: Optional[Sequence[Sequence[torch.nn.Parameter]]] = None,
                 average: bool = False, group=None, optimizer=None, time_collectives: bool = False):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.average = average
        self.buckets = [list(b) for b in (buckets if buckets is not None else stage_buckets(model))]
        self.flat: List[torch.Tensor] = []
        self._pending: List[int] = []
        self._works = []
        self._index = {}
        self._view = {}
        self.time_collectives = time_collectives   # record CUDA events around finish() (bench.py: nccl_ms_per_step)
        self._wait_events = []
        shared = None
        if optimizer is not None and hasattr(optimizer, "grad_buffer"):
            flat_all, params, offs = optimizer.grad_buffer(0)
            shared = {p: o for p, o in zip(params, offs)}
        for bi, ps in enumerate(self.buckets):
            if shared is not None:
                ps.sort(key=lambda p: shared[p])
                lo = shared[ps[0]]
                hi = lo
                for p in ps:
                    if shared[p] != hi:
                        raise ValueError("bucket parameters are not contiguous in the optimiser's flat buffer")
                    hi += (p.numel() + 3) // 4 * 4
                flat = flat_all[lo:hi]
            else:
                n = sum((p.numel() + 3) // 4 * 4 for p in ps)
                flat = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
            o = 0
            for p in ps:
                view = flat[o:o + p.numel()].view_as(p)
                p.grad = view
                self._view[p] = view
                self._index[p] = bi
                o += (p.numel() + 3) // 4 * 4
                p.register_post_accumulate_grad_hook(self._hook)
            self.flat.append(flat)
        self.reset()

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._works = []

    def zero_grad(self):
        """Use this instead of `optimizer.zero_grad()` (whose default set_to_none=True would sever the views;
        `_hook` re-attaches them if that happens anyway)."""
        for f in self.flat:
            f.zero_()
        self.reset()

    def _hook(self, p: torch.nn.Parameter):
        view = self._view[p]
        if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
            # a stock zero_grad(set_to_none=True) replaced the bucket view by a fresh tensor: copy the gradient into
            # the bucket and re-attach, so the collective never reduces a stale bucket
            view.copy_(p.grad)
            p.grad = view
        bi = self._index[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and self.world > 1:
            self._works.append(dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Block the current stream on the outstanding collectives (no host sync on NCCL)."""
        if self.world > 1:
            for bi, left in enumerate(self._pending):
                if left > 0:  # parameters that received no gradient this step
                    self._works.append(dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group,
                                                       async_op=True))
            timed = self.time_collectives and torch.cuda.is_available() and not torch.cuda.is_current_stream_capturing()
            if timed:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            for w in self._works:
                w.wait()
            if timed:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                self._wait_events.append((e0, e1))
            if self.average:
                for f in self.flat:
                    f.div_(self.world)
        self._works = []

    def exposed_wait_ms(self) -> float:
        """Mean time the compute stream spent blocked on the collectives per finish() (the part of the all-reduce
        that backward did not hide); needs time_collectives=True and a host sync."""
        if not self._wait_events:
            return 0.0
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in self._wait_events]
        self._wait_events = []
        return sum(ms) / len(ms)


class DataParallelTrainer:
    """Minimal DP training driver mirroring the step body of the reference train loop
    (/root/reference/src/train.py:38-75): zero_grad -> forward -> loss -> isfinite -> backward ->
    [all-reduce mean] -> clip -> step."""

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer: torch.optim.Optimizer,
                 max_norm_grad: Optional[float] = None):
        self.model, self.loss_fn, self.optimizer, self.max_norm_grad = model, loss_fn, optimizer, max_norm_grad
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        fused = hasattr(optimizer, "grad_buffer")
        if self.world > 1 and fused:
            # the fused optimiser consumes the reducer's buckets in place and folds 1/world into its kernel
            self.reducer = BucketedGradAllReduce(model, average=False, optimizer=optimizer)
            optimizer.grad_scale = 1.0 / self.world
        else:
            self.reducer = BucketedGradAllReduce(model, average=True) if self.world > 1 else None
        if self.world > 1:
            # replicas start identical (DDP broadcasts rank 0's state at construction)
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0)

    def step(self, data: torch.Tensor, target: torch.Tensor):
        self.model.train()
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        output = self.model(data)
        loss = self.loss_fn(output, target)
        # The reference skips the step on a non-finite loss (train.py:55-60).  Under data parallelism that decision
        # must be COLLECTIVE: a rank that returned early would launch no all-reduce and dead-lock the others, and the
        # replicas' weights would diverge.  One MIN all-reduce of a finite flag; every rank skips if any rank must.
        finite = torch.isfinite(loss.detach()).to(torch.float32).reshape(1)
        if self.world > 1:
            dist.all_reduce(finite, op=dist.ReduceOp.MIN)
        if finite.item() == 0.0:
            return loss, output
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        if self.max_norm_grad:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_norm_grad)
        self.optimizer.step()
        return loss, output
