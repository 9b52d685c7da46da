"""ORACLE tooling -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

It imports the reference's own modules (src.models.R2Plus1D, src.loss) read-only from
/root/reference with no-op stubs for packages that are not installed (pytorch_model_summary), runs
them on seeded synthetic inputs in fp32 on the CPU, and stores small summaries.  Nothing of the
reference's source is copied into this repository.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: golden vectors can only be regenerated in the build container")
    stub = types.ModuleType("pytorch_model_summary")
    stub.summary = lambda *a, **k: ""
    sys.modules.setdefault("pytorch_model_summary", stub)
    sys.path.insert(0, REF)
    from src.models.R2Plus1D import R2Plus1DClassifier  # noqa
    from src.loss import CELoss, FocalLoss, LDAMLoss  # noqa
    return R2Plus1DClassifier, FocalLoss, LDAMLoss, CELoss


def slowfast_golden():
    """E: SlowFast [1,2,2,1] (config 3): seed-42 state summary, the constructor probe's BatchNorm buffers (they are
    rounding-noise driven in the deep layers, so they are stored, not re-derived) and one small training step."""
    from src.models.slowfast import SlowFast
    from src.models.resnet import Bottleneck3D
    from src.loss import FocalLoss
    out = {}
    T, H, W, B = 20, 64, 64, 4
    torch.manual_seed(42)
    m = SlowFast((3, T, H, W), Bottleneck3D, [1, 2, 2, 1], 4, 1, 2, 1.0)
    sd = m.state_dict()
    out["keys"] = np.array(list(sd.keys()))
    out["shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])
    out["summary"] = np.stack([summarise(v) for v in sd.values()])
    bn_keys = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
    out["init_bn_keys"] = np.array(bn_keys)
    out["init_bn_values"] = np.concatenate([sd[k].numpy().reshape(-1) for k in bn_keys]).astype(np.float32)
    x, y = synthetic(B, T, H, W)
    y[0], y[1] = 0, 1
    out["y"] = y.numpy()
    w = torch.FloatTensor([0.98, 0.02])
    m.train()
    logits = m(x)
    loss = FocalLoss(weight=w, gamma=2.0)(logits, y)
    loss.backward()
    out["logits"] = logits.detach().numpy()
    out["loss"] = np.array(loss.item())
    names = [n for n, p in m.named_parameters()]
    out["grad_names"] = np.array(names)
    out["grad_norm"] = np.array([0.0 if p.grad is None else p.grad.double().norm().item() for _, p in m.named_parameters()])
    out["grad_summary"] = np.stack([summarise(p.grad if p.grad is not None else torch.zeros(1)) for _, p in m.named_parameters()])
    sd2 = m.state_dict()
    out["bn_summary"] = np.stack([summarise(sd2[k]) for k in bn_keys])
    m.eval()
    with torch.no_grad():
        out["eval_logits"] = m(x).numpy()
    np.savez_compressed(os.path.join(OUT, "slowfast_step.npz"), **out)
    print("slowfast: %d keys, loss %.6f" % (len(sd), loss.item()))


def summarise(t: torch.Tensor) -> np.ndarray:
    """[sum, sum|.|, first 4 values] in float64 -- enough to pin a tensor without storing it."""
    f = t.detach().double().reshape(-1)
    head = torch.zeros(4, dtype=torch.float64)
    head[:min(4, f.numel())] = f[:4]
    return np.concatenate([[f.sum().item(), f.abs().sum().item()], head.numpy()])


def synthetic(B, T, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (B, 3, T, H, W), generator=g).float()
    x -= torch.tensor([90.0, 98.0, 102.0]).view(1, 3, 1, 1, 1)
    y = torch.randint(0, 2, (B,), generator=g)
    return x, y


def main():
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    R2Plus1DClassifier, FocalLoss, LDAMLoss, CELoss = import_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---- A: initial weights of the benchmark model for seed 42 -------------------------------
    torch.manual_seed(42)
    model = R2Plus1DClassifier((3, 21, 128, 128), 2, [1, 2, 2, 1], False, 1.0)
    sd = model.state_dict()
    np.savez_compressed(os.path.join(OUT, "init_seed42.npz"),
                        keys=np.array(list(sd.keys())),
                        shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
                        summary=np.stack([summarise(v) for v in sd.values()]))
    print("init_seed42: %d keys, %d params" % (len(sd), sum(p.numel() for p in model.parameters())))

    # ---- B: one small training step, per alpha and loss --------------------------------------
    cls_num_list = [300, 17000]
    w_rw = 1.0 / np.array(cls_num_list)
    w_rw = torch.FloatTensor(w_rw / np.sum(w_rw))
    out = {}
    B, T, H, W = 4, 5, 32, 32
    x, y = synthetic(B, T, H, W)
    y[0], y[1] = 0, 1
    out["x_summary"] = summarise(x)
    out["y"] = y.numpy()
    for alpha in (1.0, 0.01):
        for loss_name in ("focal", "ldam", "ce"):
            torch.manual_seed(42)
            m = R2Plus1DClassifier((3, T, H, W), 2, [1, 1, 1, 1], False, alpha)
            m.train()
            if loss_name == "focal":
                lf = FocalLoss(weight=w_rw, gamma=2.0)
            elif loss_name == "ldam":
                lf = LDAMLoss(cls_num_list, max_m=0.5, weight=w_rw, s=1.0)
            else:
                lf = CELoss(weight=w_rw)
            logits = m(x)
            loss = lf(logits, y)
            loss.backward()
            tag = f"a{alpha}_{loss_name}"
            out[tag + "_logits"] = logits.detach().numpy()
            out[tag + "_loss"] = np.array(loss.item())
            names = [n for n, _ in m.named_parameters()]
            out[tag + "_grad_names"] = np.array(names)
            out[tag + "_grad_summary"] = np.stack([summarise(p.grad) for _, p in m.named_parameters()])
            out[tag + "_grad_norm"] = np.array([p.grad.double().norm().item() for _, p in m.named_parameters()])
            sdm = m.state_dict()
            bn_keys = [k for k in sdm if k.endswith("running_mean") or k.endswith("running_var")]
            out[tag + "_bn_keys"] = np.array(bn_keys)
            out[tag + "_bn_summary"] = np.stack([summarise(sdm[k]) for k in bn_keys])
            # eval-mode logits after the step's running-stat update (no optimiser step taken)
            m.eval()
            with torch.no_grad():
                out[tag + "_eval_logits"] = m(x).numpy()
            print(tag, "loss", loss.item())
    np.savez_compressed(os.path.join(OUT, "small_train_step.npz"), **out)

    # ---- C: losses on random logits ----------------------------------------------------------
    out = {}
    g = torch.Generator().manual_seed(7)
    for C_ in (2, 5):
        n = 64
        logits = (torch.randn(n, C_, generator=g) * 3).requires_grad_(True)
        target = torch.randint(0, C_, (n,), generator=g)
        counts = [300 * (i + 1) ** 2 for i in range(C_)]
        w = torch.rand(C_, generator=g) + 0.1
        out[f"c{C_}_logits"] = logits.detach().numpy()
        out[f"c{C_}_target"] = target.numpy()
        out[f"c{C_}_weight"] = w.numpy()
        out[f"c{C_}_counts"] = np.array(counts)
        cases = {
            "focal_g2": FocalLoss(weight=w, gamma=2.0),
            "focal_g0": FocalLoss(weight=w, gamma=0.0),
            "focal_g1.5": FocalLoss(weight=w, gamma=1.5),
            "ldam_s30": LDAMLoss(counts, max_m=0.5, weight=w, s=30),
            "ldam_s1": LDAMLoss(counts, max_m=0.5, weight=w, s=1.0),
            "ldam_s1_now": LDAMLoss(counts, max_m=0.5, weight=None, s=1.0),
            "ce": CELoss(weight=w),
            "ce_now": CELoss(weight=None),
        }
        for name, lf in cases.items():
            logits.grad = None
            val = lf(logits, target)
            val.backward()
            out[f"c{C_}_{name}_loss"] = np.array(val.item())
            out[f"c{C_}_{name}_grad"] = logits.grad.detach().numpy().copy()
            if name.startswith("ldam"):
                out[f"c{C_}_{name}_m"] = lf.m_list.numpy()
    np.savez_compressed(os.path.join(OUT, "loss_kat.npz"), **out)

    # ---- D: RW / DRW class weights (formulas at train_vision_network.py:312-318, src/train.py:318-329;
    #         train_DRW's helper is a closure, so the arithmetic is replayed here with the same numpy calls)
    out = {"rw": w_rw.numpy()}
    num_epoch, betas = 128, [0, 0.25, 0.5, 0.75]
    for epoch in (0, 31, 32, 63, 64, 95, 96, 127):
        idx = epoch // int(num_epoch / len(betas))
        idx = min(idx, len(betas) - 1)
        beta = betas[idx]
        eff = 1.0 - np.power(beta, cls_num_list)
        w = (1.0 - beta) / np.array(eff)
        w = w / np.sum(w) * len(cls_num_list)
        out[f"drw_e{epoch}"] = torch.FloatTensor(w).numpy()
    small = [3, 40]
    for epoch in (0, 32, 64, 96):
        beta = betas[epoch // 32]
        eff = 1.0 - np.power(beta, small)
        w = (1.0 - beta) / np.array(eff)
        out[f"drw_small_e{epoch}"] = torch.FloatTensor(w / np.sum(w) * 2).numpy()
    np.savez_compressed(os.path.join(OUT, "class_weights.npz"), **out)
    slowfast_golden()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
