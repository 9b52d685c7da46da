"""Secondary workloads of bench.py (BASELINE.json configs 3, 4, 5, the loss-kernel sweep of SURVEY 8d) and the
multi-rank parity check.  Each prints ONE JSON line in bench.py's contract; none of them is the headline metric.

    python bench.py --workload infer       # config 5: sliding-window inference over a synthetic 1024-frame shot
    python bench.py --workload slowfast    # config 3: SlowFast train step
    python bench.py --workload multimodal  # config 4: R(2+1)D + 0D transformer + GradientBlending train step (DP under torchrun)
    python bench.py --workload loss        # fused Focal / LDAM / CE kernel at N = 2^24 rows: GB/s against the HBM peak
    torchrun ... bench.py --check --gpus 2 # rank-local loss vs the oracle on that rank's shard; all-reduced gradients vs
                                           # the mean of the per-shard oracle gradients; eager and graph-captured
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MEAN = (90.0, 98.0, 102.0)
LAYER_SIZES = [1, 2, 2, 1]
CLS_NUM = [300, 17000]
FWD_GFLOP_PER_CLIP = 22.748
STEM_SPATIAL_GFLOP_PER_CLIP = 1.138   # conv1.spatio_conv (layer table, row 0): 21 frames x 54.2 MFLOP
SLOWFAST_FWD_GFLOP_PER_CLIP = 0.900


def _common(args):
    import torch
    import dp_b200  # noqa: F401
    from dp_b200 import _lib, distributed as dpd
    import bench
    rank, local_rank, world = dpd.init_distributed()
    _lib.require_device()
    dev = torch.device("cuda", local_rank)
    return torch, _lib.load(), dpd, bench, rank, local_rank, world, dev


def _sync(torch, dist, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(torch, dist, world, dev, ms):
    if world == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _finish(world):
    sys.stdout.flush()
    if world > 1:
        os._exit(0)


# ------------------------------------------------------------------------------------------------
# config 5: continuous sliding-window inference (utility.py:896-977; published: 22.53 it/s on an RTX 3090 incl. JPEG
# reads, 0.1198 s batch-1 CPU latency -- BASELINE.md section 1)
# ------------------------------------------------------------------------------------------------
def run_infer(args):
    torch, lib, dpd, bench, rank, local_rank, world, dev = _common(args)
    import torch.distributed as dist
    import dp_b200
    from dp_b200 import functional as Fn, inference
    from dp_b200.R2Plus1D import R2Plus1DClassifier

    peaks = bench.load_peaks()
    torch.manual_seed(42)
    model = R2Plus1DClassifier((3, 21, 128, 128), 2, LAYER_SIZES, False, args.alpha).to(dev).eval()
    n_frames, seq_len, dgap = 1024, 21, 3
    g = torch.Generator().manual_seed(4321)
    frames_host = torch.randint(0, 256, (n_frames, 128, 128, 3), generator=g, dtype=torch.uint8).pin_memory()
    frames = frames_host.to(dev)
    n_win = inference.num_windows(n_frames, seq_len, dgap)          # 1000
    rng = dpd.shard_range(n_win, rank, world)

    def one_pass(fr, bs, cache=True, graph=True):
        return inference.sliding_window_probs(model, fr, seq_len, dgap, batch_size=bs, window_range=rng, stem_cache=cache,
                                              use_graph=graph)

    sweep = {}
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    with dp_b200.compute_mode(args.mode, args.conv_impl):
        # parity guard inside the bench: cached-stem fused path == per-window path on a few windows
        a = inference.sliding_window_probs(model, frames, seq_len, dgap, batch_size=8, window_range=range(0, 16))
        b = inference.sliding_window_probs(model, frames, seq_len, dgap, batch_size=8, window_range=range(0, 16), stem_cache=False)
        assert (a - b).abs().max().item() < 2e-2, "stem-cache path disagrees with the per-window path"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for bs in (1, 8, 64, 256):
            reps = max(1, args.steps if bs >= 8 else min(args.steps, 2))
            for _ in range(max(1, min(args.warmup, 3 if bs >= 8 else 1))):
                one_pass(frames, bs)
            _sync(torch, dist, world)
            if bs == 256:
                sampler.mark()
            l0 = lib.dp_launch_count()
            e0.record()
            for _ in range(reps):
                p = one_pass(frames, bs)
            e1.record()
            _sync(torch, dist, world)
            ms = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1)) / reps
            sweep[bs] = {"windows_per_s": round(n_win / (ms / 1e3), 1), "ms_per_shot": round(ms, 3),
                         "launches_per_shot": int((lib.dp_launch_count() - l0) // reps)}
        # batch 1 with every kernel issued from Python (no CUDA-graph replay): the launch-bound regime of the reference loop
        one_pass(frames, 1, graph=False)
        _sync(torch, dist, world)
        e0.record()
        one_pass(frames, 1, graph=False)
        e1.record()
        _sync(torch, dist, world)
        ms_b1_eager = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1))
        # without the per-frame stem cache (what batching alone buys)
        for _ in range(2):
            one_pass(frames, 256, cache=False)
        _sync(torch, dist, world)
        e0.record()
        for _ in range(args.steps):
            one_pass(frames, 256, cache=False)
        e1.record()
        _sync(torch, dist, world)
        ms_nocache = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1)) / args.steps
        # end to end: the shot's uint8 frames start in pinned HOST memory, probabilities are read back
        bs = 256
        e0.record()
        for _ in range(args.steps):
            fr = frames_host.to(dev, non_blocking=True)
            probs_host = one_pass(fr, bs).cpu()
        e1.record()
        _sync(torch, dist, world)
        ms_e2e = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1)) / args.steps
        # batch-1 latency of ONE window through the public call (lines up with measure_computation_time, utility.py:1201)
        lat = []
        x1 = torch.zeros((1, 3, 21, 128, 128), device=dev)
        with torch.no_grad():
            for i in range(16 + 3):
                t0 = time.time()
                out = model(x1)
                out.cpu()
                if i >= 3:
                    lat.append(time.time() - t0)
        # per-kernel events of one batch-256 pass
        kern = {}
        if rank == 0:
            Fn.PROFILER = Fn.KernelProfiler()
            torch.cuda._sleep(int(6e7))
            one_pass(frames, 256, graph=False)
            kern = Fn.PROFILER.summary()
            Fn.PROFILER = None
        clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        _finish(world)
        return
    best = sweep[256]
    value = best["windows_per_s"]
    conv = kern.get("tc_gather_gemm", {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
    roofline = None
    if conv["launches"]:
        tf = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        roofline = {"kernel": "tc_gather_gemm (eval: conv + BatchNorm(running stats) + LeakyReLU [+ residual] in one kernel)",
                    "bound": "tensor", "achieved": round(tf, 2), "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": round(tf / peaks["bf16_tflops_sustained"], 4), "traffic": None,
                    "avg_launch_ms": round(conv["ms"] / conv["launches"], 4), "launches_per_shot": conv["launches"],
                    "ms_per_shot": round(conv["ms"], 3), "hbm_gbs_same_kernel": round(conv["bytes"] / (conv["ms"] * 1e-3) / 1e9, 1),
                    "peak_source": peaks["source"] + " (sustained)"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_infer_baseline()
    import numpy as np
    line = {
        "metric": "r2plus1d_sliding_window_windows_per_sec", "value": value, "unit": "windows/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["ms_per_shot"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": round(sweep[1]["windows_per_s"] / 22.53, 2),
        "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "config 5: sliding-window inference (utility.py:896-977) over a synthetic 1024-frame uint8 shot, seq_len 21, "
                               "dist 3 -> 1000 windows, eval mode; one step = the whole shot; windows sharded over ranks, no collective",
                   "batch_sweep": {str(k): v for k, v in sweep.items()}, "value_is": "batch 256 per GPU",
                   "batch1_without_graph_replay": {"windows_per_s": round(n_win / (ms_b1_eager / 1e3), 1), "ms_per_shot": round(ms_b1_eager, 3)},
                   "without_stem_frame_cache": {"windows_per_s": round(n_win / (ms_nocache / 1e3), 1), "ms_per_shot": round(ms_nocache, 3)},
                   "batch1_latency_s": {"mean": round(float(np.mean(lat)), 5), "std": round(float(np.std(lat)), 5), "samples": len(lat),
                                        "how": "model(x) on a (1,3,21,128,128) fp32 clip + .cpu(), time.time(), as measure_computation_time "
                                               "(utility.py:1201); published 0.1198 s +- 0.0292 on the authors' CPU"},
                   "vs_baseline_is": "batch-1 windows/s / the published 22.53 it/s (RTX 3090, includes 21 JPEG reads per window; "
                                     "BASELINE.md section 1) -- different hardware, orientation only",
                   "l2": "activations of a 256-window batch (> 1 GB) exceed the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": round(n_win / (ms_e2e / 1e3), 1), "unit": "windows/s", "h2d_bytes_per_step": frames_host.numel(),
                "d2h_bytes_per_step": int(probs_host.numel() * 4), "ms_per_step": round(ms_e2e, 3),
                "api": "inference.sliding_window_probs(model, frames_u8) with the shot in pinned host memory, probabilities read back"},
        "gpu_launches": best["launches_per_shot"] * args.steps,
        "simt_launches": int(lib.dp_simt_launch_count()), "simt_fallbacks": int(lib.dp_simt_fallback_count()),
        "roofline": roofline,
        "kernels": {k: {"ms_per_shot": round(v["ms"], 3), "launches": v["launches"]} for k, v in kern.items()},
        "effective_tflops_reference_work": round(FWD_GFLOP_PER_CLIP * value / 1e3, 1),
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    _finish(world)


def cpu_infer_baseline(budget_s: float = 25.0):
    """The reference's loop on the host cores: batch 1, eval mode, one window per forward (oracle port)."""
    import torch
    from oracle import r2plus1d_port as port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = port.clone_state(port.init_state(LAYER_SIZES, 2, seed=42), requires_grad=False)
    x = torch.zeros(1, 3, 21, 128, 128)
    ts = []
    with torch.no_grad():
        port.classifier_forward(st, x, LAYER_SIZES, 1.0, training=False)
        t_end = time.perf_counter() + budget_s
        while len(ts) < 16 and time.perf_counter() < t_end:
            t0 = time.perf_counter()
            port.classifier_forward(st, x, LAYER_SIZES, 1.0, training=False)
            ts.append(time.perf_counter() - t0)
    mean = sum(ts) / len(ts)
    return {"value": round(1.0 / mean, 3), "unit": "windows/s", "cores": cores, "kind": "port",
            "batch1_latency_s": round(mean, 5),
            "sample": f"{len(ts)} batch-1 eval forwards of the same model (fp32, torch CPU primitives the reference calls); "
                      f"the reference loop runs one window per forward (utility.py:936-949)"}


# ------------------------------------------------------------------------------------------------
# generic train-step harness for the secondary models
# ------------------------------------------------------------------------------------------------
def _train_harness(args, name, metric, build, make_batch, workload, flops_per_clip=None, cpu=None):
    torch, lib, dpd, bench, rank, local_rank, world, dev = _common(args)
    import torch.distributed as dist
    import dp_b200
    from dp_b200 import functional as Fn
    from dp_b200.optim import FusedClipAdamW

    peaks = bench.load_peaks()
    B = args.batch
    torch.manual_seed(42)
    model, loss_fn, forward = build(torch, dev)
    model = model.to(dev).train()
    opt = FusedClipAdamW(model.parameters(), lr=2e-4, max_norm=1.0, capturable=True)
    reducer = None
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
        reducer = dpd.BucketedGradAllReduce(model, average=False, optimizer=opt, time_collectives=True)
        opt.grad_scale = 1.0 / world
    host_batches = [make_batch(torch, B, 1234 + rank + 100 * i) for i in range(2)]
    host_batches = [tuple(t.pin_memory() for t in hb) for hb in host_batches]
    dev_batches = [tuple(t.to(dev) for t in hb) for hb in host_batches]

    def step(batch):
        *xs, y = batch
        if reducer is not None:
            reducer.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        out = forward(model, *xs)
        loss = loss_fn(*out, y) if isinstance(out, tuple) else loss_fn(out, y)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    graphed = None

    def run_step(batch):
        if graphed is None:
            return step(batch)
        *xs, y = batch
        return graphed.step(tuple(xs) if len(xs) > 1 else xs[0], y)[0]

    with dp_b200.compute_mode(args.mode, args.conv_impl):
        for i in range(max(3, args.warmup)):
            loss = step(dev_batches[i % 2])
        _sync(torch, dist, world)
        assert torch.isfinite(loss).item(), "non-finite loss in warm-up"
        if args.graph:
            from dp_b200.graph import GraphedTrainStep
            loss = None
            # anything that still references the eager autograd graph (the GB model caches its latents as attributes) keeps
            # AccumulateGrad nodes alive on the legacy stream, which a capture on a side stream may not depend on
            for mod in model.modules():
                for attr in ("vis_latent", "ts_latent"):
                    if getattr(mod, attr, None) is not None:
                        setattr(mod, attr, None)
            kw = dict(pre_backward=reducer.zero_grad, post_backward=reducer.finish) if reducer is not None else {}
            *xs, y0 = dev_batches[0]
            graphed = GraphedTrainStep(model, loss_fn, opt, tuple(xs) if len(xs) > 1 else xs[0], y0, warmup=1,
                                       forward=lambda *a: forward(model, *a), **kw)
            for i in range(max(3, args.warmup)):
                loss = run_step(dev_batches[i % 2])
            _sync(torch, dist, world)
            assert torch.isfinite(loss).item(), "non-finite loss after graph capture"
        sampler.mark()
        l0, sl0, sf0 = lib.dp_launch_count(), lib.dp_simt_launch_count(), lib.dp_simt_fallback_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _sync(torch, dist, world)
        e0.record()
        for i in range(args.steps):
            loss = run_step(dev_batches[i % 2])
        e1.record()
        _sync(torch, dist, world)
        ms_dev = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1))
        launches = int(lib.dp_launch_count() - l0) if graphed is None else graphed.launches_per_step * args.steps
        simt = int(lib.dp_simt_launch_count() - sl0)
        fallbacks = int(lib.dp_simt_fallback_count() - sf0)
        final_loss = float(loss.item())
        # end to end: batch from pinned host memory every step, loss read back
        _sync(torch, dist, world)
        e0.record()
        for i in range(args.steps):
            if graphed is not None:      # the replay's static buffers are filled straight from pinned host memory
                run_step(host_batches[i % 2]).item()
            else:
                step(tuple(t.to(dev, non_blocking=True) for t in host_batches[i % 2])).item()
        e1.record()
        _sync(torch, dist, world)
        ms_e2e = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None
        kern = {}
        if reducer is not None:
            reducer.exposed_wait_ms()
        if rank == 0:
            Fn.PROFILER = Fn.KernelProfiler()
        for i in range(args.profile_steps):
            torch.cuda._sleep(int(6e7))
            step(dev_batches[i % 2])
        nccl_ms = reducer.exposed_wait_ms() if (reducer is not None and args.profile_steps) else None
        if rank == 0:
            kern = Fn.PROFILER.summary()
            Fn.PROFILER = None
        _sync(torch, dist, world)
    if rank != 0:
        _finish(world)
        return
    clips = B * world * args.steps
    fam = {}
    for k, d in kern.items():
        ms = d["ms"] / max(1, args.profile_steps)
        fam[k] = {"ms_per_step": round(ms, 4), "launches_per_step": d["launches"] // max(1, args.profile_steps),
                  "tflops": round(d["flops"] / max(1, args.profile_steps) / (ms * 1e-3) / 1e12, 2) if ms > 0 else 0.0,
                  "gbs": round(d["bytes"] / max(1, args.profile_steps) / (ms * 1e-3) / 1e9, 1) if ms > 0 else 0.0}
    roofline = None
    if fam:
        top = max(fam, key=lambda k: fam[k]["ms_per_step"])
        r = fam[top]
        if kern[top]["flops"] > 0:
            roofline = {"kernel": top, "bound": "tensor", "achieved": r["tflops"], "peak": peaks["bf16_tflops_sustained"],
                        "unit": "TFLOP/s", "frac": round(r["tflops"] / peaks["bf16_tflops_sustained"], 4), "traffic": None,
                        "avg_launch_ms": round(r["ms_per_step"] / max(1, r["launches_per_step"]), 4)}
        else:
            roofline = {"kernel": top, "bound": "hbm", "achieved": r["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": round(r["gbs"] / peaks["hbm_gbs"], 4), "traffic": None,
                        "avg_launch_ms": round(r["ms_per_step"] / max(1, r["launches_per_step"]), 4)}
    h2d = sum(t.numel() * t.element_size() for t in host_batches[0])
    line = {
        "metric": metric, "value": round(clips / (ms_dev / 1e3), 2), "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(ms_dev / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload, "global_batch": B * world, "parallelism": f"dp{world}" if world > 1 else "single",
                   "launch": "one CUDA graph replay per step" if graphed is not None else "eager (one Python call per kernel)",
                   "final_loss": final_loss,
                   "l2": "activations of one step exceed the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": round(clips / (ms_e2e / 1e3), 2), "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3), "api": "model(...) / loss / backward / optimizer.step() from pinned host memory"},
        "gpu_launches": launches, "simt_launches": simt, "simt_fallbacks": fallbacks,
        "nccl_exposed_ms_per_step": None if nccl_ms is None else round(nccl_ms, 4),
        "roofline": roofline, "kernels": fam,
        "conv_fwd_gflop_per_clip": flops_per_clip,
        "cpu_baseline": cpu() if (cpu is not None and world == 1 and not args.no_cpu_baseline) else None,
    }
    print(json.dumps(line), flush=True)
    _finish(world)


def run_slowfast(args):
    """Config 3: SlowFast [1,2,2,1] on (3,20,128,128) clips (reference slowfast.py:163-173, resnet.py:172-273)."""
    def build(torch, dev):
        from dp_b200.slowfast import Bottleneck3D, SlowFast
        from dp_b200.loss import FocalLoss
        import dp_b200
        m = SlowFast((3, 20, 128, 128), Bottleneck3D, LAYER_SIZES, 4, 1, 2, args.alpha)
        w = dp_b200.rw_class_weights(CLS_NUM)
        return m, FocalLoss(weight=w.to(dev), gamma=2.0), (lambda model, x: model(x))

    def make_batch(torch, B, seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.randint(0, 256, (B, 3, 20, 128, 128), generator=g, dtype=torch.uint8).float()
        x -= torch.tensor(MEAN).view(1, 3, 1, 1, 1)
        y = torch.randint(0, 2, (B,), generator=g)
        y[0], y[1 % B] = 0, 1
        return x, y

    _train_harness(args, "slowfast", "slowfast_train_clips_per_sec", build, make_batch,
                   f"config 3: SlowFast((3,20,128,128), Bottleneck3D, [1,2,2,1], alpha_slowfast=4) train step: fwd + Focal + bwd + "
                   f"clip+AdamW, batch {args.batch}/GPU, 74 convs on the tcgen05 kernels, native SE/Swish/MaxPool/concat kernels",
                   flops_per_clip=SLOWFAST_FWD_GFLOP_PER_CLIP)


def run_multimodal(args):
    """Config 4: R(2+1)D video encoder + 0D transformer + three heads, GradientBlending loss (weights .1/.4/.5,
    train_multimodal.py:375-385), data parallel under torchrun."""
    def build(torch, dev):
        import dp_b200
        from dp_b200.MultiModal import GradientBlending, MultiModalR2Plus1D_GB
        from dp_b200.loss import FocalLoss
        args_v = {"layer_sizes": LAYER_SIZES, "alpha": args.alpha}
        args_t = dict(n_features=18, kernel_size=5, feature_dims=128, max_len=21, n_layers=2, n_heads=8, dim_feedforward=256,
                      dropout=0.1)
        m = MultiModalR2Plus1D_GB(2, args_v, args_t, use_stream="multi-GB")
        w = dp_b200.rw_class_weights(CLS_NUM).to(dev)
        gb = GradientBlending(FocalLoss(weight=w, gamma=2.0), FocalLoss(weight=w, gamma=2.0), FocalLoss(weight=w, gamma=2.0),
                              vis_weight=0.1, ts_weight=0.4, vis_ts_weight=0.5)
        return m, gb, (lambda model, xv, xt: model(xv, xt))

    def make_batch(torch, B, seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.randint(0, 256, (B, 3, 21, 128, 128), generator=g, dtype=torch.uint8).float()
        x -= torch.tensor(MEAN).view(1, 3, 1, 1, 1)
        ts = torch.randn(B, 21, 18, generator=g)
        y = torch.randint(0, 2, (B,), generator=g)
        y[0], y[1 % B] = 0, 1
        return x, ts, y

    _train_harness(args, "multimodal", "multimodal_gb_train_clips_per_sec", build, make_batch,
                   f"config 4: MultiModalR2Plus1D_GB (R(2+1)D [1,2,2,1] video branch + 0D TransformerEncoder(18 features, 21 steps) + fusion, "
                   f"three heads) train step with GradientBlending(Focal x3, .1/.4/.5) + clip+AdamW, batch {args.batch}/GPU",
                   flops_per_clip=FWD_GFLOP_PER_CLIP)


# ------------------------------------------------------------------------------------------------
# loss-kernel sweep ("measure on N = 2^24 rows for a GB/s figure; in situ it is launch-bound")
# ------------------------------------------------------------------------------------------------
def run_loss(args):
    torch, lib, dpd, bench, rank, local_rank, world, dev = _common(args)
    from dp_b200 import _lib as L, functional as Fn
    import ctypes as C
    peaks = bench.load_peaks()
    rows = {}
    N, Cc = 1 << 24, 2
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(N, Cc, generator=g).to(dev)
    target = torch.randint(0, Cc, (N,), generator=g).to(dev)
    w = torch.tensor([0.98, 0.02], device=dev)
    margins = torch.tensor([0.5, 0.18], device=dev)
    res = torch.empty(2, dtype=torch.float32, device=dev)
    dl = torch.empty_like(logits)
    ws = torch.zeros(int(lib.dp_loss_workspace(N)), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sp = L.stream_ptr()
    algo_bytes = N * (Cc * 4 + 8 + Cc * 4)        # logits in, int64 target in, dlogits out: 24 B/row at C = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for kind, name, gamma, s in ((L.LOSS_FOCAL, "focal", 2.0, 1.0), (L.LOSS_LDAM, "ldam", 0.0, 30.0), (L.LOSS_CE, "ce", 0.0, 1.0)):
        ts = []
        for i in range(args.warmup + args.steps):
            flush.zero_()                          # 256 MB write between timed launches: L2 (126 MB) holds nothing of the inputs
            e0.record()
            L.check(lib.dp_loss_fwd_bwd(kind, logits.data_ptr(), target.data_ptr(), w.data_ptr(),
                                        margins.data_ptr() if kind == L.LOSS_LDAM else None, gamma, s, N, Cc,
                                        res.data_ptr(), dl.data_ptr(), ws.data_ptr(), sp), "dp_loss_fwd_bwd")
            e1.record()
            torch.cuda.synchronize()
            if i >= args.warmup:
                ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        rows[name] = {"ms": round(ms, 4), "gbs": round(algo_bytes / (ms * 1e-3) / 1e9, 1), "value": float(res[0].item())}
    # in situ: the real shape (B = 64 rows) -- launch-bound
    lg, tg = logits[:64].contiguous(), target[:64].contiguous()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        L.check(lib.dp_loss_fwd_bwd(L.LOSS_FOCAL, lg.data_ptr(), tg.data_ptr(), w.data_ptr(), None, 2.0, 1.0, 64, Cc, res.data_ptr(),
                                    dl.data_ptr(), ws.data_ptr(), sp), "dp_loss_fwd_bwd")
    e1.record()
    torch.cuda.synchronize()
    in_situ_us = e0.elapsed_time(e1) * 10.0
    top = rows["focal"]
    line = {"metric": "loss_kernel_hbm_gbs", "value": top["gbs"], "unit": "GB/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": top["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "fused loss forward+backward kernel (csrc/loss.cu) on N = 2^24 rows x 2 classes, one launch per step: "
                                   "Focal(gamma=2, weighted) is the value; LDAM(s=30) and CE beside it",
                       "rows": N, "algorithmic_bytes_per_row": algo_bytes // N, "l2": "256 MB flush write between timed launches",
                       "kinds": rows, "in_situ_us_per_launch_at_64_rows": round(in_situ_us, 2)},
            "roofline": {"kernel": "loss_kernel", "bound": "hbm", "achieved": top["gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(top["gbs"] / peaks["hbm_gbs"], 4), "traffic": None, "peak_source": peaks["source"] + " (burst: timed alone)",
                         "frac_of_nominal_8TBs": round(top["gbs"] / 8000.0, 4)},
            "gpu_launches": 3 * args.steps}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# multi-rank parity on real NCCL (SURVEY 8e "parity method")
# ------------------------------------------------------------------------------------------------
def run_check(args):
    """Every rank: (1) its rank-local logits / loss against the oracle port run on THAT rank's shard (fp32 validation
    mode: 1e-4); (2) the all-reduced gradients against the mean over ranks of the per-shard oracle gradients
    (all-gathered), for the eager reducer and for the graph-captured step; (3) the replicas' weights stay bit-identical
    after optimiser steps.  Prints one JSON line with the worst deviations and exits non-zero on failure."""
    torch, lib, dpd, bench, rank, local_rank, world, dev = _common(args)
    import torch.distributed as dist
    import dp_b200
    from dp_b200.R2Plus1D import R2Plus1DClassifier
    from dp_b200.loss import FocalLoss
    from dp_b200.optim import FusedClipAdamW
    from dp_b200.graph import GraphedTrainStep
    from oracle import r2plus1d_port as port           # checker only

    assert world > 1, "run under torchrun with --gpus >= 2"
    layer_sizes, alpha, T, H, W, Bl = [1, 1, 1, 1], 0.01, 9, 64, 64, 4
    res = {"world": world}
    ok = True
    w = dp_b200.rw_class_weights(CLS_NUM)
    x, y = port.structured_clips(Bl, T, H, W, seed=100 + rank)     # this rank's shard
    y[0], y[1] = 0, 1
    torch.manual_seed(42)
    state = port.init_state(layer_sizes, 2, seed=42)

    # the oracle on this rank's shard (CPU fp32) and the mean of the per-shard gradients
    st = port.clone_state(state)
    ref_logits, ref_loss, ref_grads = port.train_step(st, x, y, layer_sizes, alpha, "focal", w)
    names = sorted(ref_grads)
    flat_ref = torch.cat([ref_grads[k].reshape(-1) for k in names]).to(dev)
    gathered = [torch.empty_like(flat_ref) for _ in range(world)]
    dist.all_gather(gathered, flat_ref)
    mean_ref = torch.stack(gathered).mean(0)

    def make(mode_fused=True):
        m = R2Plus1DClassifier((3, T, H, W), 2, layer_sizes, False, alpha)
        m.load_state_dict(state)
        m = m.to(dev).train()
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=0)
        return m

    def flat_grads(m):
        d = dict(m.named_parameters())
        return torch.cat([d[k].grad.reshape(-1) for k in names])

    lf = FocalLoss(weight=w.to(dev), gamma=2.0)
    for mode, tol_l, tol_g in (("fp32", 1e-4, 2e-2), ("bf16", 0.3, None)):
        with dp_b200.compute_mode(mode):
            m = make()
            red = dpd.BucketedGradAllReduce(m, average=True)
            red.zero_grad()
            logits = m(x.to(dev))
            loss = lf(logits, y.to(dev))
            loss.backward()
            red.finish()
            torch.cuda.synchronize()
            e_logit = ((logits.detach().cpu() - ref_logits).abs().max() / ref_logits.abs().max()).item()
            e_loss = abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
            g = flat_grads(m)
            e_grad = ((g - mean_ref).norm() / mean_ref.norm()).item()
            # all ranks must hold the same reduced gradient
            gs = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(gs, g)
            same = all(torch.equal(gs[0], t) for t in gs)
            res[mode] = {"rank_local_logits_rel": e_logit, "rank_local_loss_rel": e_loss,
                         "allreduced_grad_rel_l2_vs_mean_of_shard_oracle_grads": e_grad, "identical_on_all_ranks": same}
            ok &= e_logit < tol_l and e_loss < tol_l and same and (tol_g is None or e_grad < tol_g)
            del red, m
    # graph-captured collective path: FusedClipAdamW consuming the reducer's buckets in place, NCCL inside the graph;
    # compare three replayed steps with three eager steps of the same trainer construction, and the replicas' weights
    with dp_b200.compute_mode("bf16"):
        def trainer():
            m = make()
            opt = FusedClipAdamW(m.parameters(), lr=1e-3, max_norm=1.0, capturable=True)
            red = dpd.BucketedGradAllReduce(m, average=False, optimizer=opt)
            opt.grad_scale = 1.0 / world
            return m, opt, red
        xd, yd = x.to(dev), y.to(dev)
        m1, o1, r1 = trainer()
        eager = []
        for _ in range(4):
            r1.zero_grad()
            l = lf(m1(xd), yd)
            l.backward()
            r1.finish()
            o1.step()
            eager.append(l.item())
        del l
        m2, o2, r2 = trainer()
        gs = GraphedTrainStep(m2, lf, o2, xd, yd, warmup=1, pre_backward=r2.zero_grad, post_backward=r2.finish)
        graphed = [gs.step(xd, yd)[0].item() for _ in range(3)]
        torch.cuda.synchronize()
        e_traj = max(abs(a - b) / max(abs(a), 1e-6) for a, b in zip(eager[1:], graphed))
        p1 = torch.cat([p.detach().reshape(-1) for p in m1.parameters()])
        p2 = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
        e_w = ((p1 - p2).abs().max() / p1.abs().max()).item()
        ps = [torch.empty_like(p2) for _ in range(world)]
        dist.all_gather(ps, p2)
        replicas_equal = all(torch.equal(ps[0], t) for t in ps)
        res["graph_captured"] = {"loss_trajectory_rel_vs_eager": e_traj, "weights_rel_vs_eager": e_w,
                                 "replica_weights_bit_identical": replicas_equal, "eager_losses": eager, "graphed_losses": graphed}
        ok &= e_traj < 5e-3 and e_w < 5e-3 and replicas_equal
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ok"] = bool(flag.item() == 1.0)
    if rank == 0:
        print(json.dumps({"check": "multi_rank_parity", **res}), flush=True)
    sys.stdout.flush()
    os._exit(0 if res["ok"] else 1)
