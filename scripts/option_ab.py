"""A/B of load-time options (dp_set_option) on the whole BASELINE training step inside ONE process: the model, inputs and
optimiser are built once; for every option set the step is re-captured into a CUDA graph (the options are read when the
kernels are launched, i.e. at capture), warmed up and timed with CUDA events over `--steps` replays, `--reps` times.

    python scripts/option_ab.py base "bn_sweep=0" rev "bn_sweep=7" ...        (pairs: label, comma-separated options)

Prints one line per option set: label, best and median ms/step, clips/s, final loss."""
import argparse
import gc
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pairs", nargs="+")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    assert len(args.pairs) % 2 == 0, "label/options pairs"

    import torch
    import dp_b200
    from dp_b200 import _lib
    from dp_b200.R2Plus1D import R2Plus1DClassifier
    from dp_b200.loss import FocalLoss
    from dp_b200.optim import FusedClipAdamW
    from dp_b200.graph import GraphedTrainStep
    import bench

    _lib.require_device()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = R2Plus1DClassifier(bench.CLIP, 2, bench.LAYER_SIZES, False, 1.0).to(dev).train()
    weights = dp_b200.drw_class_weights(40, 128, dp_b200.drw_betas(0.25), bench.CLS_NUM)
    loss_fn = FocalLoss(weight=weights.to(dev), gamma=2.0)
    opt = FusedClipAdamW(model.parameters(), lr=2e-4, max_norm=1.0, capturable=True)
    g = torch.Generator().manual_seed(1234)
    xs = []
    for _ in range(2):
        x = torch.randint(0, 256, (args.batch, *bench.CLIP), generator=g, dtype=torch.uint8).float()
        x -= torch.tensor([90.0, 98.0, 102.0]).view(1, 3, 1, 1, 1)
        xs.append(x.to(dev))
    y = torch.randint(0, 2, (args.batch,), generator=g)
    y[0], y[1] = 0, 1
    y = y.to(dev)

    lines = []
    with dp_b200.compute_mode("bf16", "auto"):
        for _ in range(3):      # eager warm-up (packs, workspaces, planner caches)
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(model(xs[0]), y)
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        loss = None
        defaults = {}
        for label, opts in zip(args.pairs[0::2], args.pairs[1::2]):
            changed = []
            for item in filter(None, opts.split(",")):
                name, val = item.split("=")
                if name not in defaults:
                    defaults[name] = lib.dp_get_option(name.encode())
                assert lib.dp_set_option(name.encode(), int(val)) == 0, item
                changed.append(name)
            graphed = GraphedTrainStep(model, loss_fn, opt, xs[0], y, warmup=1)
            for i in range(3):
                loss = graphed.step(xs[i % 2], y)[0]
            torch.cuda.synchronize()
            times = []
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(args.steps):
                    loss = graphed.step(xs[i % 2], y)[0]
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) / args.steps)
            line = (f"{label:14s} best {min(times):7.3f} median {statistics.median(times):7.3f} ms/step  "
                    f"{args.batch / min(times) * 1e3:8.1f} clips/s  loss {float(loss.item()):.6f}  [{opts}]")
            print(line, flush=True)
            lines.append(line)
            for name in changed:
                lib.dp_set_option(name.encode(), defaults[name])
            del graphed
            loss = None
            gc.collect()
            torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
