"""ORACLE -- test infrastructure.  Generates tests/golden/convergence_ref.npz: the loss / accuracy trajectory of the
UNMODIFIED reference (its own R2Plus1DClassifier and FocalLoss imported from /root/reference, stock
torch.optim.AdamW + clip_grad_norm_, the step body of /root/reference/src/train.py:38-75; fp32, CPU) on the
learnable synthetic task of oracle/synth_task.py, at the BASELINE model
R2Plus1DClassifier((3,21,128,128),2,[1,2,2,1],alpha=1.0) built under torch.manual_seed(42).  300 steps x 16 clips
take ~20 min on 8 cores, which is why the trajectory is a committed fixture: tests/test_gpu_convergence.py replays
the same batches through the CUDA path (bf16 product mode and fp32 validation mode) and compares the curves.
With --storage bf16 the oracle port runs the same trajectory with bf16-STORED activations (r2plus1d_port.py
storage emulation) from the same initial state: what bf16 storage alone does to the reference's algorithm.

Run in the build container only (needs /root/reference):

    python oracle/make_convergence_golden.py [--steps 300] [--batch 16] [--storage fp32|bf16] [--out PATH]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2plus1d_port as port          # noqa: E402
from oracle.synth_task import task_batch          # noqa: E402

LAYER_SIZES = [1, 2, 2, 1]
ALPHA = 1.0
LR = 2e-4          # the reference trainer's default learning rate (train_vision_network.py:72)
MAX_NORM = 1.0     # src/train.py:63-64


def reference_model(size):
    from oracle.make_golden import import_reference
    R2Plus1DClassifier, FocalLoss, _, _ = import_reference()
    torch.manual_seed(42)
    return R2Plus1DClassifier((3, 21, size, size), 2, LAYER_SIZES, False, ALPHA), FocalLoss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--storage", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--heldout", type=int, default=64, help="held-out clips scored in eval mode after training")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "convergence_ref.npz"))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    model, FocalLoss = reference_model(args.size)
    w = torch.ones(2)
    if args.storage == "fp32":          # the reference itself
        model.train()
        loss_fn = FocalLoss(weight=w, gamma=2.0)
        params = list(model.parameters())
        opt = torch.optim.AdamW(params, lr=LR)
        fwd = model
    else:                               # the port with bf16-stored activations, same initial state
        st = port.clone_state({k: v for k, v in model.state_dict().items()})
        params = [v for v in st.values() if v.requires_grad]
        opt = torch.optim.AdamW(params, lr=LR)
        fwd = lambda x: port.classifier_forward(st, x, LAYER_SIZES, ALPHA, True, storage="bf16")   # noqa: E731
        loss_fn = lambda o, t: port.focal_loss(o, t, w, 2.0)                                        # noqa: E731
    losses, accs = [], []
    t0 = time.time()
    for s in range(args.steps):
        x, y = task_batch(s, args.batch, 21, args.size, args.size)
        opt.zero_grad(set_to_none=True)
        logits = fwd(x)
        loss = loss_fn(logits, y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, MAX_NORM)
        opt.step()
        losses.append(float(loss))
        accs.append(float((logits.argmax(1) == y).float().mean()))
        if s % 10 == 0:
            print(f"step {s} loss {losses[-1]:.4f} acc {accs[-1]:.2f} ({time.time() - t0:.0f}s)", flush=True)
    # held-out clips (steps >= 100000 of the same generator family), eval mode, final weights
    held_logits = None
    if args.storage == "fp32" and args.heldout > 0:
        model.eval()
        with torch.no_grad():
            held_logits = torch.cat([model(task_batch(100000 + i, 16, 21, args.size, args.size)[0])
                                     for i in range(args.heldout // 16)]).numpy()
        held_y = torch.cat([task_batch(100000 + i, 16, 21, args.size, args.size)[1] for i in range(args.heldout // 16)]).numpy()
    extra = {} if held_logits is None else {"heldout_logits": held_logits, "heldout_y": held_y}
    np.savez(args.out, loss=np.asarray(losses, np.float64), acc=np.asarray(accs, np.float64), **extra,
             steps=args.steps, batch=args.batch, size=args.size, lr=LR, max_norm=MAX_NORM, alpha=ALPHA,
             storage=args.storage, source="reference" if args.storage == "fp32" else "port-bf16-storage")
    print("wrote", args.out)


if __name__ == "__main__":
    main()
