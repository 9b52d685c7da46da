"""Convergence A/B on a learnable synthetic task (test infrastructure).

The BASELINE model R2Plus1DClassifier((3,21,128,128),2,[1,2,2,1],alpha=1.0), seed-42 initial state, is trained for
`steps` optimiser steps of `batch` clips of oracle/synth_task.py (label carried by a temporal brightness collapse) with
the reference's step body (/root/reference/src/train.py:38-75: Focal loss, clip_grad_norm_(1.0), torch.optim.AdamW
lr 2e-4) in three arms that see IDENTICAL batches:

    bf16   the CUDA path, product mode (bf16 storage, fp32 accumulate, tcgen05 kernels)
    fp32   the CUDA path, fp32 validation mode
    port   the oracle port (the reference's algorithm: torch conv3d / batch_norm / ..., fp32, TF32 off) on the same GPU

and, when tests/golden/convergence_ref.npz is present, against the trajectory of the UNMODIFIED reference on the CPU
(oracle/make_convergence_golden.py).  Afterwards `heldout` unseen clips are scored in eval mode: thresholded
disruption labels not(softmax[:,0] > 0.5) (src/evaluate.py:56-57) of the bf16 and fp32 CUDA paths against the port
running the SAME weights, and each arm's own held-out accuracy.

    python tests/convergence_ab.py [--steps 300] [--batch 16] [--heldout 1024] [--out gpurun_out/r2_convergence.json]
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LS, ALPHA, LR, MAX_NORM = [1, 2, 2, 1], 1.0, 2e-4, 1.0
MEAN = (90.0, 98.0, 102.0)


def _batches(n, B, size, first=0):
    """uint8 (B,3,T,H,W) clips on the GPU + labels; the generator is CPU-side and deterministic in (step)."""
    from oracle.synth_task import task_batch
    xs, ys = [], []
    mean = torch.tensor(MEAN).view(1, 3, 1, 1, 1)
    for s in range(n):
        x, y = task_batch(first + s, B, 21, size, size)
        xs.append((x + mean).round().to(torch.uint8).cuda())
        ys.append(y.cuda())
    return xs, ys


def _clip(xu8):
    return xu8.float() - torch.tensor(MEAN, device=xu8.device).view(1, 3, 1, 1, 1)


def run_ab(steps=300, batch=16, heldout=1024, size=128, log=print):
    import dp_b200
    from dp_b200.R2Plus1D import R2Plus1DClassifier
    from dp_b200.loss import FocalLoss
    from oracle import r2plus1d_port as port      # checker arm

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    torch.manual_seed(42)
    init = R2Plus1DClassifier((3, 21, size, size), 2, LS, False, ALPHA)
    state = {k: v.clone() for k, v in init.state_dict().items()}
    w = torch.ones(2)
    t0 = time.time()
    xs, ys = _batches(steps, batch, size)
    hx, hy = _batches(heldout // 16, 16, size, first=100000)
    log(f"generated {steps} training and {len(hx)} held-out batches in {time.time() - t0:.0f}s")
    res = {"steps": steps, "batch": batch, "heldout": len(hx) * 16, "lr": LR, "max_norm": MAX_NORM, "alpha": ALPHA, "size": size}

    def train_cuda(mode):
        m = R2Plus1DClassifier((3, 21, size, size), 2, LS, False, ALPHA)
        m.load_state_dict(state)
        m = m.to(dev).train()
        lf = FocalLoss(weight=w.to(dev), gamma=2.0)
        opt = torch.optim.AdamW(m.parameters(), lr=LR)
        losses, accs = [], []
        with dp_b200.compute_mode(mode):
            for s in range(steps):
                opt.zero_grad()
                out = m(_clip(xs[s]))
                loss = lf(out, ys[s])
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), MAX_NORM)
                opt.step()
                losses.append(loss.detach())
                accs.append((out.argmax(1) == ys[s]).float().mean())
        return m, torch.stack(losses).cpu().numpy().astype(np.float64), torch.stack(accs).cpu().numpy().astype(np.float64)

    def train_port():
        st = {k: v.to(dev) for k, v in port.clone_state(state).items()}
        st = {k: (v.detach().requires_grad_(True) if k.endswith(("weight", "bias")) else v.detach()) for k, v in st.items()}
        params = [v for v in st.values() if v.requires_grad]
        opt = torch.optim.AdamW(params, lr=LR)
        wd = w.to(dev)
        losses, accs = [], []
        for s in range(steps):
            opt.zero_grad()
            out = port.classifier_forward(st, _clip(xs[s]), LS, ALPHA, True)
            loss = port.focal_loss(out, ys[s], wd, 2.0)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, MAX_NORM)
            opt.step()
            losses.append(loss.detach())
            accs.append((out.argmax(1) == ys[s]).float().mean())
        return st, torch.stack(losses).cpu().numpy().astype(np.float64), torch.stack(accs).cpu().numpy().astype(np.float64)

    def probs_cuda(m, mode):
        m.eval()
        with dp_b200.compute_mode(mode), torch.no_grad():
            return torch.cat([torch.softmax(m(_clip(x)), 1)[:, 0] for x in hx])

    def probs_port(st):
        with torch.no_grad():
            return torch.cat([torch.softmax(port.classifier_forward(st, _clip(x), LS, ALPHA, False), 1)[:, 0] for x in hx])

    t0 = time.time()
    m16, l16, a16 = train_cuda("bf16")
    log(f"bf16 arm: {time.time() - t0:.0f}s, last-50 loss {l16[-50:].mean():.4f} acc {a16[-50:].mean():.3f}")
    t0 = time.time()
    m32, l32, a32 = train_cuda("fp32")
    log(f"fp32 arm: {time.time() - t0:.0f}s, last-50 loss {l32[-50:].mean():.4f} acc {a32[-50:].mean():.3f}")
    t0 = time.time()
    stp, lp, ap_ = train_port()
    log(f"port arm: {time.time() - t0:.0f}s, last-50 loss {lp[-50:].mean():.4f} acc {ap_[-50:].mean():.3f}")
    res["loss"] = {"bf16": l16.tolist(), "fp32": l32.tolist(), "port": lp.tolist()}
    res["acc"] = {"bf16": a16.tolist(), "fp32": a32.tolist(), "port": ap_.tolist()}
    gold = os.path.join(ROOT, "tests", "golden", "convergence_ref.npz")
    ref = None
    if os.path.exists(gold):
        g = np.load(gold)
        if int(g["steps"]) >= steps and int(g["batch"]) == batch and int(g["size"]) == size:
            ref = g["loss"][:steps].astype(np.float64)
            res["loss"]["reference_cpu"] = ref.tolist()
            res["acc"]["reference_cpu"] = g["acc"][:steps].tolist()
    # window means of the loss curves
    edges = [e for e in (0, 10, 25, 50, 100, 200, 300, 100000) if e < steps] + [steps]
    wins = list(zip(edges[:-1], edges[1:]))
    res["windows"] = [list(wn) for wn in wins]
    res["window_mean_loss"] = {k: [float(np.mean(v[a:b])) for a, b in wins]
                               for k, v in (("bf16", l16), ("fp32", l32), ("port", lp)) + ((("reference_cpu", ref),) if ref is not None else ())}
    # held-out labels: same weights (the port arm's final state), three implementations
    hy_all = torch.cat(hy)
    m_same16 = R2Plus1DClassifier((3, 21, size, size), 2, LS, False, ALPHA)
    m_same16.load_state_dict({k: v.detach().cpu() for k, v in stp.items()})
    m_same16 = m_same16.to(dev)
    p_port = probs_port(stp)
    p_16 = probs_cuda(m_same16, "bf16")
    p_32 = probs_cuda(m_same16, "fp32")
    lab = lambda p: ~(p > 0.5)        # noqa: E731  evaluate.py:56-57: pred = not(P(class 0) > 0.5)
    res["heldout_same_weights"] = {
        "label_agreement_bf16_vs_port": float((lab(p_16) == lab(p_port)).float().mean()),
        "label_agreement_fp32_vs_port": float((lab(p_32) == lab(p_port)).float().mean()),
        "max_abs_dP_bf16_vs_port": float((p_16 - p_port).abs().max()),
        "max_abs_dP_fp32_vs_port": float((p_32 - p_port).abs().max()),
        "min_margin_port": float((p_port - 0.5).abs().min()),
        "accuracy_port": float((lab(p_port).long() == hy_all).float().mean()),
    }
    res["heldout_own_weights_accuracy"] = {
        "bf16": float((lab(probs_cuda(m16, "bf16")).long() == hy_all).float().mean()),
        "fp32": float((lab(probs_cuda(m32, "fp32")).long() == hy_all).float().mean()),
        "port": res["heldout_same_weights"]["accuracy_port"],
    }
    if ref is not None and "heldout_logits" in np.load(gold).files:
        g = np.load(gold)
        n = g["heldout_logits"].shape[0]
        pr = torch.softmax(torch.from_numpy(g["heldout_logits"]), 1)[:, 0]
        res["heldout_own_weights_accuracy"]["reference_cpu_first%d" % n] = float(((~(pr > 0.5)).long() == torch.from_numpy(g["heldout_y"])).float().mean())
    return res


def check(res, log=print):
    """The stated band: every arm learns the task, the window-mean loss curves agree, thresholded labels agree."""
    wm = res["window_mean_loss"]
    ok = True
    for k in ("bf16", "fp32"):
        for i, (a, b) in enumerate(zip(wm[k], wm["port"])):
            band = 0.35 * max(a, b) + 0.03      # 35 % relative + 0.03 absolute on each window mean (sum-reduced Focal loss of 16 clips)
            good = abs(a - b) <= band
            log(f"  window {res['windows'][i]}: {k} {a:.4f} port {b:.4f} {'ok' if good else 'OUT OF BAND'}")
            ok &= good
    hs = res["heldout_same_weights"]
    ok &= hs["label_agreement_bf16_vs_port"] >= 0.999 and hs["label_agreement_fp32_vs_port"] >= 0.999
    own = res["heldout_own_weights_accuracy"]
    ok &= min(own["bf16"], own["fp32"], own["port"]) >= 0.97
    for k in ("bf16", "fp32", "port"):
        ok &= float(np.mean(res["acc"][k][-100:])) >= 0.97
    return ok


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--heldout", type=int, default=1024)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_convergence.json"))
    a = ap.parse_args()
    r = run_ab(a.steps, a.batch, a.heldout)
    r["within_band"] = bool(check(r))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(r, open(a.out, "w"))
    print(json.dumps({k: v for k, v in r.items() if k not in ("loss", "acc")}, indent=1))
