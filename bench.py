#!/usr/bin/env python
"""bench.py -- R(2+1)D training throughput on B200 (BASELINE.json metric: train clips/sec).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference's CPU implementation on host cores

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
    """Average DRAM bytes per launch of the dominant kernel family from the committed `ncu --set full` capture
    (profiles/*_prof_*_raw.csv, dram__bytes_read.sum + dram__bytes_write.sum); None if no capture is present."""
    import csv
    import glob
    tag = {"tc_gather_gemm": "gather", "tc_wgrad": "wgrad"}.get(kernel_family, "bn")
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_prof_{tag}_raw.csv")))
    if not files:
        return None, None
    rows = list(csv.reader(open(files[-1])))
    if len(rows) < 3:
        return None, None
    h, units = rows[0], rows[1]
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for r in rows[2:]:
        try:
            v = 0.0
            for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = h.index(col)
                v += float(r[i]) * unit_scale.get(units[i], 1.0)
            tot += v
            n += 1
        except (ValueError, IndexError):
            continue
    return (tot / n if n else None), os.path.basename(files[-1])


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
def cpu_reference_run(steps: int, warmup: int, budget_s: float = 200.0):
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
    per_step_budget = budget_s / max(1, steps + warmup)
    while B > 1 and t_first * 0.9 > per_step_budget:
        B //= 2
        t_first = one(B)
    for _ in range(max(0, warmup - 1)):
        one(B)
    ts = [one(B) for _ in range(steps)]
    total = sum(ts)
    return {"clips_per_s": B * steps / total, "ms_per_step": 1e3 * total / steps, "B": B, "cores": cores,
            "steps": steps}


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
        "cpu_baseline": {"value": r["clips_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
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
    opt = FusedClipAdamW(model.parameters(), lr=2e-4, max_norm=1.0, capturable=True)
    reducer = None
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)
        opt._ensure_flat(0, opt.param_groups[0])          # re-point params into the flat bucket first
        reducer = dpd.BucketedGradAllReduce(model, average=False)
        opt.grad_scale = 1.0 / world                      # DDP-mean semantics folded into the optimiser kernel

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

    with dp_b200.compute_mode("bf16", args.conv_impl):
        # nvidia-smi takes a few hundred ms to produce its first sample: start it before the warm-up, report from mark()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        # ---- warm-up (eager), then capture the whole step into one CUDA graph and warm that up too ----
        for i in range(max(3, args.warmup)):
            loss = step(x_dev[i % n_host], y_dev)
        barrier()
        assert torch.isfinite(loss).item(), "non-finite loss in warm-up"
        if args.graph:
            from dp_b200.graph import GraphedTrainStep
            loss = None      # drop the eager autograd graph (its AccumulateGrad nodes sit on the default stream)
            if reducer is not None:    # data parallel: bucket zeroing and the all-reduce waits are part of the graph
                graphed = GraphedTrainStep(model, loss_fn, opt, x_dev[0], y_dev, warmup=1,
                                           pre_backward=reducer.zero_grad, post_backward=reducer.finish)
            else:
                graphed = GraphedTrainStep(model, loss_fn, opt, x_dev[0], y_dev, warmup=1)
            for i in range(max(3, args.warmup)):
                loss = run_step(x_dev[i % n_host], y_dev)
            barrier()
            assert torch.isfinite(loss).item(), "non-finite loss after graph capture"

        # ---- timed region A: inputs resident in HBM ----
        sampler.mark()
        l0 = lib.dp_launch_count()
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
        copy_stream = torch.cuda.Stream(device=dev)
        stage = [torch.empty_like(x_dev[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        ysrc = y_host.pin_memory()

        def issue_copy(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                stage[s].copy_(host[i % n_host], non_blocking=True)
                ready[s].record(copy_stream)

        cur = torch.cuda.current_stream()
        for s in range(2):
            freed[s].record(cur)
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
            loss = run_step(stage[s], yb)
            freed[s].record(cur)
            t_e2e.append(loss.item())                       # D2H read of the step's loss, every step
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions
        h2d = host[0].numel() * 4 + ysrc.numel() * 8
        d2h = 4

        # ---- per-kernel CUDA-event timing of the same step (roofline leg) ----
        kern = {}
        if rank == 0:
            Fn.PROFILER = Fn.KernelProfiler()
        for i in range(args.profile_steps):       # every rank steps (the step holds collectives); rank 0 records
            # gate the stream behind a ~30 ms spin so the host has queued the step before the GPU starts it:
            # the per-kernel events then bracket GPU time, not host launch latency
            torch.cuda._sleep(int(6e7))
            step(x_dev[i % n_host], y_dev)
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
        cpu = {"value": round(r["clips_per_s"], 3), "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"B={r['B']} clips per step x 3 steps (after 1 warm-up) of the same model and step, fp32, "
                         f"torch CPU primitives the reference calls (oracle/r2plus1d_port.py)"}

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"R2Plus1DClassifier((3,21,128,128),2,[1,2,2,1],alpha={alpha}) train step: fwd + Focal(gamma=2, DRW "
                               f"weights) + bwd + clip(1.0)+AdamW, batch {B}/GPU, bf16 storage / fp32 accumulate",
                   "global_batch": B * world, "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": "inputs larger than L2 (264 MB of clips and >8 GB of activations per step vs 126 MB L2)",
                   "conv_impl": args.conv_impl, "final_loss": final_loss,
                   "launch": "one CUDA graph replay per step" if graphed is not None else "eager (one Python call per kernel)"},
        "clocks": clocks,
        "e2e": None if args.no_e2e else {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 3),
                "api": "model(x) / loss_fn / loss.backward() / optimizer.step() on fp32 NCDHW clips from pinned host memory"},
        "gpu_launches": launches,
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
