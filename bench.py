#!/usr/bin/env python
"""bench.py -- R(2+1)D training throughput on B200 (BASELINE.json metric: train clips/sec).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference's CPU implementation on host cores
    python bench.py --workload infer|slowfast|multimodal|loss   # configs 5 / 3 / 4 of BASELINE.json, loss-kernel sweep
    python bench.py --caller stock                           # the UNCHANGED caller protocol of src/train.py:38-75
    python bench.py --mode fp32                              # the fp32 validation mode (CUDA-core kernels)
    python bench.py --check --gpus 2                         # multi-rank parity on real NCCL (bench_extra.py)

One "step" = one pass of the hot path over one batch of synthetic clips:
    zero_grad -> forward (32 conv+BN+LeakyReLU layers, pool, head) -> Focal loss (DRW class weights)
    -> backward (dgrad + wgrad + BN backward) -> [DP: gradient all-reduce, mean] -> clip(1.0) + AdamW
(the step body of /root/reference/src/train.py:38-75) on the model of BASELINE.json configs[1]:
R2Plus1DClassifier((3,21,128,128), 2, [1,2,2,1]) at batch 64 per GPU, bf16 storage / fp32 accumulate.

Prints ONE JSON line (rank 0).  `value` = clips/s with inputs resident in HBM; `e2e` = the same step through the
public nn.Module API with the clips in pinned HOST memory (fp32 NCDHW as the reference's DataLoader hands them,
H2D inside the timed region, loss read back every step).  `roofline` describes the dominant kernel family from
per-kernel CUDA-event timing of the same steps; `cpu_baseline` is the oracle port of the reference (same torch
CPU primitives the reference calls) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "r2plus1d_train_clips_per_sec"
UNIT = "clips/s"
LAYER_SIZES = [1, 2, 2, 1]
CLIP = (3, 21, 128, 128)
CLS_NUM = [300, 17000]
# algorithmic conv work per clip
FWD_GFLOP_PER_CLIP = 22.748
TRAIN_GFLOP_PER_CLIP = 67.11   # fwd + dgrad + wgrad without the stem's unused dgrad


def measured_traffic(kernel_family: str):
    """Mean DRAM bytes per launch of a kernel family over ALL its launches of one training step, from the committed
    ncu pass (profiles/**/*_dram_per_launch.json, written by scripts/summarize_dram.py from
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`): the same set of launches the
    algorithmic bytes per launch are averaged over.  (None, None) if no capture is present."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "**", "*_dram_per_launch.json"), recursive=True),
                   key=os.path.basename)
    if not files:
        return None, None
    try:
        d = json.load(open(files[-1]))
        fam = d["families"].get(kernel_family)
        if not fam:
            return None, None
        return fam["dram_bytes_per_launch"], f"{os.path.basename(files[-1])}: {fam['launches']} launches of one step"
    except (OSError, ValueError, KeyError):
        return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Samples before this point (process start-up, warm-up) are not reported."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows = self.rows[self.first:]
        if len(rows) < 2:          # a timed region shorter than two sampling periods: include the warm-up steps (same work)
            rows = self.rows[max(0, self.first - 5):]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle port: the same torch CPU
# primitives in the same order; /root/reference does not exist on the GPU box and has no installable package)
# ------------------------------------------------------------------------------------------------
def cpu_model_name() -> str:
    """`lscpu` model string of the host; /proc/cpuinfo when lscpu is missing."""
    try:
        out = subprocess.run(["lscpu"], capture_output=True, text=True, timeout=10).stdout
        for l in out.splitlines():
            if l.lower().startswith("model name"):
                return l.split(":", 1)[1].strip()
    except (OSError, subprocess.SubprocessError):
        pass
    try:
        for l in open("/proc/cpuinfo"):
            if l.lower().startswith("model name"):
                return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_run(steps: int, warmup: int, budget_s: float = 200.0, anomaly_steps: int = 2):
    import torch
    from oracle import r2plus1d_port as port

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    alpha = 1.0
    torch.manual_seed(42)
    st = port.clone_state(port.init_state(LAYER_SIZES, 2, seed=42))
    params = [v for v in st.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=2e-4)
    w = port.drw_class_weights(40, 128, [0, 0.25, 0.5, 0.75], CLS_NUM)
    B = 8

    def one(Bc):
        x, y = port.synthetic_clips(Bc, *CLIP[1:])
        y[0], y[1 % Bc] = 0, 1
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        logits = port.classifier_forward(st, x, LAYER_SIZES, alpha, training=True)
        loss = port.focal_loss(logits, y, w, 2.0)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        float(loss.detach())
        return time.perf_counter() - t0

    t_first = one(B)                       # also warms the allocator / thread pool
    per_step_budget = budget_s / max(1, steps + warmup + anomaly_steps)
    while B > 1 and t_first * 0.9 > per_step_budget:
        B //= 2
        t_first = one(B)
    for _ in range(max(0, warmup - 1)):
        one(B)
    ts = [one(B) for _ in range(steps)]
    total = sum(ts)
    out = {"clips_per_s": B * steps / total, "ms_per_step": 1e3 * total / steps, "B": B, "cores": cores,
           "steps": steps, "cpu_model": cpu_model_name(),
           "clips_per_s_median": B / sorted(ts)[len(ts) // 2], "clips_per_s_best": B / min(ts)}
    if anomaly_steps > 0:
        # as shipped: src/train.py:15 switches autograd anomaly detection on globally at import
        with torch.autograd.set_detect_anomaly(True):
            ta = [one(B) for _ in range(anomaly_steps)]
        out["clips_per_s_anomaly_on"] = B * len(ta) / sum(ta)
    return out


def cpu_baseline_block(r, sample: str):
    return {"value": round(r["clips_per_s"], 3), "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample,
            "cpu_model": r["cpu_model"], "median": round(r["clips_per_s_median"], 3), "best": round(r["clips_per_s_best"], 3),
            "anomaly_detection_on": None if "clips_per_s_anomaly_on" not in r else round(r["clips_per_s_anomaly_on"], 3),
            "anomaly_note": "value/median/best: anomaly detection off (fair); anomaly_detection_on: as shipped (src/train.py:15)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    sample = f"B={r['B']} clips per step x {r['steps']} steps of the configs[1] model (fwd+Focal+bwd+clip+AdamW, fp32)"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["clips_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "R2Plus1DClassifier((3,21,128,128),2,[1,2,2,1]) train step, Focal+DRW, CPU host cores",
                   "batch_per_step": r["B"]},
        "cpu_baseline": dict(cpu_baseline_block(r, sample), value=r["clips_per_s"]),
        "e2e": {"value": r["clips_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dp_b200
    from dp_b200 import _lib, distributed as dpd, functional as Fn
    from dp_b200.R2Plus1D import R2Plus1DClassifier
    from dp_b200.loss import FocalLoss
    from dp_b200.optim import FusedClipAdamW

    if os.environ.get("DP_BENCH_FAULT"):      # debugging aid: dump every thread's Python stack after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["DP_BENCH_FAULT"]), exit=True)
    rank, local_rank, world = dpd.init_distributed()
    _lib.require_device()
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    peaks = load_peaks()
    B = args.batch
    alpha = args.alpha

    torch.manual_seed(42)
    model = R2Plus1DClassifier(CLIP, 2, LAYER_SIZES, False, alpha).to(dev).train()
    weights = dp_b200.drw_class_weights(40, 128, dp_b200.drw_betas(0.25), CLS_NUM)   # DRW epoch 40 of 128: beta=.25
    loss_fn = FocalLoss(weight=weights.to(dev), gamma=2.0)
    stock = args.caller == "stock"
    if stock:      # what an unmodified src/train.py gets: torch.optim.AdamW + clip_grad_norm_ + its three host syncs
        opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    else:
        opt = FusedClipAdamW(model.parameters(), lr=2e-4, max_norm=1.0, capturable=True)
    reducer = None
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
        if stock:
            reducer = dpd.BucketedGradAllReduce(model, average=True, time_collectives=True)
        else:
            # the reducer's buckets ARE contiguous slices of the optimiser's flat gradient buffer: no second buffer, no copy
            reducer = dpd.BucketedGradAllReduce(model, average=False, optimizer=opt, time_collectives=True)
            opt.grad_scale = 1.0 / world                  # DDP-mean semantics folded into the optimiser kernel

    # synthetic clips of the survey's distribution (uint8 grey levels minus BGR mean), seed 1234 + rank
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 2
    host = []
    for _ in range(n_host):
        x = torch.randint(0, 256, (B, *CLIP), generator=g, dtype=torch.uint8).float()
        x -= torch.tensor([90.0, 98.0, 102.0]).view(1, 3, 1, 1, 1)
        host.append(x.pin_memory())
    y_host = torch.randint(0, 2, (B,), generator=g)
    y_host[0], y_host[1] = 0, 1
    x_dev = [h.to(dev) for h in host]
    y_dev = y_host.to(dev)
    torch.cuda.synchronize()

    def step(x, y):
        if reducer is not None:
            reducer.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        out = model(x)
        loss = loss_fn(out, y)
        if stock:
            # /root/reference/src/train.py:52-75 verbatim: isfinite (host sync 1), backward, clip, step, loss.item()
            # (sync 2), argmax accuracy .item() (sync 3)
            if not torch.isfinite(loss):
                return loss
            loss.backward()
            if reducer is not None:
                reducer.finish()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            step.train_loss += loss.item()
            pred = torch.nn.functional.softmax(out, dim=1).max(1, keepdim=True)[1]
            step.train_acc += pred.eq(y.view_as(pred)).sum().item()
            return loss
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    step.train_loss, step.train_acc = 0.0, 0

    MEAN = (90.0, 98.0, 102.0)

    def step_u8(frames, y):
        """The same step fed through the uint8 input boundary (n3): (B,T,H,W,3) uint8 BGR frames, mean subtraction and
        layout change fused into the stem's input pack on the device (dataset.py:104-110,201-205)."""
        if reducer is not None:
            reducer.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        loss = loss_fn(model(frames, mean_bgr=MEAN), y)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    graphed = None

    def run_step(x, y):
        """One training step: CUDA-graph replay when the step was captured, else the eager calls."""
        if graphed is not None:
            return graphed.step(x, y)[0]
        return step(x, y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    use_graph = args.caller == "graph" and args.graph
    u8 = args.e2e_input == "u8" and args.mode == "bf16" and not stock
    with dp_b200.compute_mode(args.mode, args.conv_impl):
        # nvidia-smi takes a few hundred ms to produce its first sample: start it before the warm-up, report from mark()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        # ---- warm-up (eager), then capture the whole step into one CUDA graph and warm that up too ----
        for i in range(max(3, args.warmup)):
            loss = step(x_dev[i % n_host], y_dev)
        barrier()
        assert torch.isfinite(loss).item(), "non-finite loss in warm-up"
        graphed_u8 = None
        frames_host = frames_dev = None
        if u8:      # the e2e leg ships uint8 frames: (B,T,H,W,3), 1.03 MB per clip instead of 4.13 MB
            gq = torch.Generator().manual_seed(4321 + rank)
            frames_host = [torch.randint(0, 256, (B, CLIP[1], CLIP[2], CLIP[3], 3), generator=gq, dtype=torch.uint8).pin_memory()
                           for _ in range(n_host)]
            frames_dev = frames_host[0].to(dev)
            loss = step_u8(frames_dev, y_dev)
        if use_graph:
            from dp_b200.graph import GraphedTrainStep
            loss = None      # drop the eager autograd graph (its AccumulateGrad nodes sit on the default stream)
            kw = {}
            if reducer is not None:    # data parallel: bucket zeroing and the all-reduce waits are part of the graph
                kw = dict(pre_backward=reducer.zero_grad, post_backward=reducer.finish)
            graphed = GraphedTrainStep(model, loss_fn, opt, x_dev[0], y_dev, warmup=1, **kw)
            if u8:
                graphed_u8 = GraphedTrainStep(model, loss_fn, opt, frames_dev, y_dev, warmup=1,
                                              forward=lambda f: model(f, mean_bgr=MEAN), **kw)
            for i in range(max(3, args.warmup)):
                loss = run_step(x_dev[i % n_host], y_dev)
            barrier()
            assert torch.isfinite(loss).item(), "non-finite loss after graph capture"

        # ---- timed region A: inputs resident in HBM ----
        sampler.mark()
        l0, sl0, sf0 = lib.dp_launch_count(), lib.dp_simt_launch_count(), lib.dp_simt_fallback_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            loss = run_step(x_dev[i % n_host], y_dev)
        e1.record()
        barrier()
        ms_dev = max_over_ranks(e0.elapsed_time(e1))
        launches = int(lib.dp_launch_count() - l0) if graphed is None else graphed.launches_per_step * args.steps
        final_loss = float(loss.item())

        # ---- timed region B: end to end from pinned host memory (double-buffered H2D on a copy stream) ----
        e2e_steps = 0 if args.no_e2e else args.steps
        src_host = frames_host if u8 else host
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [torch.empty_like(frames_dev if u8 else x_dev[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        ysrc = y_host.pin_memory()

        def run_e2e_step(xb, yb):
            if u8:
                return graphed_u8.step(xb, yb)[0] if graphed_u8 is not None else step_u8(xb, yb)
            return run_step(xb, yb)

        def issue_copy(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                stage[s].copy_(src_host[i % n_host], non_blocking=True)
                ready[s].record(copy_stream)

        cur = torch.cuda.current_stream()
        for s in range(2):
            freed[s].record(cur)
        for i in range(2 if e2e_steps else 0):     # the e2e graph's own warm-up (untimed)
            run_e2e_step(frames_dev if u8 else x_dev[0], y_dev)
        barrier()
        t_e2e = []
        e0.record()
        if e2e_steps:
            issue_copy(0)
        for i in range(e2e_steps):
            s = i % 2
            if i + 1 < e2e_steps:
                issue_copy(i + 1)
            cur.wait_event(ready[s])
            yb = ysrc.to(dev, non_blocking=True)
            loss = run_e2e_step(stage[s], yb)
            freed[s].record(cur)
            t_e2e.append(loss.item())                       # D2H read of the step's loss, every step
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions
        h2d = src_host[0].numel() * src_host[0].element_size() + ysrc.numel() * 8
        d2h = 4
        simt_launches = int(lib.dp_simt_launch_count() - sl0)
        simt_fallbacks = int(lib.dp_simt_fallback_count() - sf0)

        if args.nvtx_step:    # one extra EAGER step between cudaProfilerStart/Stop (and inside the NVTX range "dp_step"):
            # `ncu --profile-from-start off` sees exactly one step, forward AND backward (the backward launches come from
            # autograd's worker thread, which a thread-bound NVTX push/pop range filter silently drops)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            torch.cuda.nvtx.range_push("dp_step")
            step(x_dev[0], y_dev)
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
            torch.cuda.profiler.stop()

        # ---- per-kernel CUDA-event timing of the same step (roofline leg) ----
        kern = {}
        if reducer is not None:
            reducer.exposed_wait_ms()      # drop the events of the eager warm-up steps
        if rank == 0:
            Fn.PROFILER = Fn.KernelProfiler()
        for i in range(args.profile_steps):       # every rank steps (the step holds collectives); rank 0 records
            # gate the stream behind a ~30 ms spin so the host has queued the step before the GPU starts it:
            # the per-kernel events then bracket GPU time, not host launch latency
            torch.cuda._sleep(int(6e7))
            step(x_dev[i % n_host], y_dev)
        nccl_exposed_ms = reducer.exposed_wait_ms() if (reducer is not None and args.profile_steps) else None
        if rank == 0:
            kern = Fn.PROFILER.summary()
            if os.environ.get("DP_BENCH_DUMP"):     # per-launch list of the last profiled step (development aid)
                recs = Fn.PROFILER.records
                n = len(recs) // max(1, args.profile_steps)
                with open(os.environ["DP_BENCH_DUMP"], "w") as f:
                    for fam, fl, nb, a, b in recs[-n:]:
                        ms = a.elapsed_time(b)
                        f.write(f"{fam:20s} {ms*1e3:9.1f} us  {fl/ms/1e9 if ms > 0 else 0:8.1f} TF/s {nb/ms/1e6 if ms > 0 else 0:8.1f} GB/s  {nb/1e6:9.1f} MB\n")
            Fn.PROFILER = None
        barrier()

    if rank != 0:
        # NCCL teardown blocks while a captured graph still holds collectives: leave without destructors
        sys.stdout.flush()
        os._exit(0)
    clips = B * world * args.steps
    value = clips / (ms_dev / 1e3)
    e2e_value = clips / (ms_e2e / 1e3) if not args.no_e2e else None

    fam_rows = {}
    for fam, d in kern.items():
        ms = d["ms"] / max(1, args.profile_steps)
        fam_rows[fam] = {"ms_per_step": round(ms, 4), "launches_per_step": d["launches"] // max(1, args.profile_steps),
                         "tflops": round(d["flops"] / max(1, args.profile_steps) / (ms * 1e-3) / 1e12, 2) if ms > 0 else 0.0,
                         "gbs": round(d["bytes"] / max(1, args.profile_steps) / (ms * 1e-3) / 1e9, 1) if ms > 0 else 0.0}
    roofline = None
    if fam_rows:
        top = max(fam_rows, key=lambda k: fam_rows[k]["ms_per_step"])
        r = fam_rows[top]
        n_l = max(1, r["launches_per_step"])
        traffic, traffic_src = measured_traffic(top)
        if kern[top]["flops"] > 0:
            peak = peaks["bf16_tflops_sustained"]
            roofline = {"kernel": top, "bound": "tensor", "achieved": r["tflops"], "peak": peak, "unit": "TFLOP/s",
                        "frac": round(r["tflops"] / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                        "traffic_over_algorithmic": None if not traffic else round(traffic / (kern[top]["bytes"] / kern[top]["launches"]), 3),
                        "algorithmic_bytes_per_launch": round(kern[top]["bytes"] / kern[top]["launches"], 1),
                        "peak_source": peaks["source"] + " (sustained: kernel timed inside a long step)",
                        "avg_launch_ms": round(r["ms_per_step"] / n_l, 4),
                        "algorithmic_gflop_per_launch": round(kern[top]["flops"] / kern[top]["launches"] / 1e9, 2),
                        "hbm_gbs_same_kernel": r["gbs"], "hbm_frac_same_kernel": round(r["gbs"] / peaks["hbm_gbs"], 4)}
        else:
            peak = peaks["hbm_gbs"]
            roofline = {"kernel": top, "bound": "hbm", "achieved": r["gbs"], "peak": peak, "unit": "GB/s",
                        "frac": round(r["gbs"] / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peaks["source"],
                        "avg_launch_ms": round(r["ms_per_step"] / n_l, 4),
                        "algorithmic_mb_per_launch": round(kern[top]["bytes"] / kern[top]["launches"] / 1e6, 2)}
    conv_ms = sum(v["ms_per_step"] for k, v in fam_rows.items() if "gemm" in k or "wgrad" in k)
    step_ms = ms_dev / args.steps

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=3, warmup=1, budget_s=40.0)
        cpu = cpu_baseline_block(r, f"B={r['B']} clips per step x 3 steps (after 1 warm-up) of the same model and step, fp32, "
                                    f"torch CPU primitives the reference calls (oracle/r2plus1d_port.py)")

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"R2Plus1DClassifier((3,21,128,128),2,[1,2,2,1],alpha={alpha}) train step: fwd + Focal(gamma=2, DRW "
                               f"weights) + bwd + clip(1.0)+AdamW, batch {B}/GPU, "
                               + ("bf16 storage / fp32 accumulate" if args.mode == "bf16" else "fp32 validation mode (CUDA-core kernels)"),
                   "global_batch": B * world, "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": "inputs larger than L2 (264 MB of clips and >8 GB of activations per step vs 126 MB L2)",
                   "conv_impl": args.conv_impl, "final_loss": final_loss,
                   "launch": "one CUDA graph replay per step" if graphed is not None else "eager (one Python call per kernel)",
                   "caller": {"graph": "opt-in: FusedClipAdamW + GraphedTrainStep", "eager": "opt-in: FusedClipAdamW, eager launches",
                              "stock": "unchanged caller (src/train.py:38-75): torch.optim.AdamW + clip_grad_norm_ + 3 host syncs per step"}[args.caller if (args.graph or args.caller != "graph") else "eager"]},
        "clocks": clocks,
        "e2e": None if args.no_e2e else {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 3),
                "api": ("model(frames_u8, mean_bgr) / loss_fn / loss.backward() / optimizer.step() on (B,T,H,W,3) uint8 frames from pinned "
                        "host memory (mean subtraction + layout on the device)") if u8 else
                       "model(x) / loss_fn / loss.backward() / optimizer.step() on fp32 NCDHW clips from pinned host memory"},
        "gpu_launches": launches,
        "simt_launches": simt_launches, "simt_fallbacks": simt_fallbacks,
        "nccl_exposed_ms_per_step": None if nccl_exposed_ms is None else round(nccl_exposed_ms, 4),
        "roofline": roofline,
        "kernels": fam_rows,
        "conv": {"ms_per_step": round(conv_ms, 3),
                 "tflops_algorithmic": round(TRAIN_GFLOP_PER_CLIP * B / max(conv_ms, 1e-9), 2),
                 "frac_of_bf16_peak": round(TRAIN_GFLOP_PER_CLIP * B / max(conv_ms, 1e-9) / peaks["bf16_tflops_sustained"], 4),
                 "share_of_step": round(conv_ms / step_ms, 3)},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step")
    ap.add_argument("--alpha", type=float, default=1.0, help="LeakyReLU slope (trainer default 1.0)")
    ap.add_argument("--conv-impl", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--profile-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="issue every kernel from Python instead of replaying the captured step")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-memory leg (used for short ncu runs)")
    ap.add_argument("--workload", default="train", choices=["train", "infer", "slowfast", "multimodal", "loss"],
                    help="train = BASELINE.json configs[1] (the headline); infer / slowfast / multimodal = configs 5 / 3 / 4; "
                         "loss = the N = 2^24 sweep of the fused loss kernel")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"], help="bf16 product mode or the fp32 validation mode")
    ap.add_argument("--caller", default="graph", choices=["graph", "eager", "stock"],
                    help="graph/eager: opt-in FusedClipAdamW (captured / eager); stock: the unchanged src/train.py step body")
    ap.add_argument("--e2e-input", default="u8", choices=["u8", "f32"],
                    help="what the end-to-end leg ships from pinned host memory: uint8 frames (n3 boundary) or fp32 NCDHW clips")
    ap.add_argument("--nvtx-step", action="store_true", help="run one extra eager step inside the NVTX range 'dp_step' (ncu captures)")
    ap.add_argument("--check", action="store_true", help="multi-rank parity check on real NCCL instead of a timing run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.caller != "graph":
        args.graph = False
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.check:
        import bench_extra
        return bench_extra.run_check(args)
    if args.workload != "train":
        import bench_extra
        return getattr(bench_extra, "run_" + args.workload)(args)
    run_ours(args)


if __name__ == "__main__":
    main()
