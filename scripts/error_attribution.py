"""Per-layer error attribution of the bf16 product mode at the BASELINE config.

For the model of BASELINE.json configs[1] -- R2Plus1DClassifier((3,21,128,128),2,[1,2,2,1],alpha) -- at batch B (default
64), in training mode, on the GPU box:

  * oracle: the port (oracle/r2plus1d_port.py) run in FLOAT64 on the same device (torch eager), every activation tapped;
  * CUDA path, bf16 product mode, module by module (forward hooks on every Conv3dBlock and residual block):
      - ACCUMULATED error of every activation against the fp64 oracle (relative L2);
      - LOCAL error of every Conv3dBlock: the block alone, fed the oracle's own input (what one layer adds);
  * features / logits / loss errors, and the head's amplification (oracle head applied to OUR features);
  * weight-gradient error of every conv against fp64 (fused path, no hooks), median / p90;
  * what-if: the last stage (conv5), pool and head in fp32 storage (the < 1 % of bytes the verdict asks about);
  * the same numbers for the oracle with bf16-STORED activations (fp32 math): what storage alone does.

Writes profiles/<tag>_error_attribution.md.   Usage: python scripts/error_attribution.py [--batch 64] [--alpha 1.0]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dp_b200                                        # noqa: E402
from dp_b200 import functional as Fn                  # noqa: E402
from dp_b200.R2Plus1D import R2Plus1DClassifier       # noqa: E402
from dp_b200.loss import FocalLoss                    # noqa: E402
from oracle import r2plus1d_port as port              # noqa: E402  (checker)

LS = [1, 2, 2, 1]
DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


class CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return x.to(dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.src), None


def forward_fp32_tail(model, x):
    """conv1..conv4 in bf16, conv5 + pool + head on fp32-stored activations."""
    from dp_b200.R2Plus1D import _call
    enc = model.res2plus1d
    stem = enc.conv1.spatio_conv
    geom = stem._stem_geom(x)
    h = _call(enc.conv1.temporal_conv, stem.forward_stem(x, geom, None))
    for stage in (enc.conv2, enc.conv3, enc.conv4):
        h = _call(stage, h)
    c = h._dp_c
    h = Fn.tag(CastFn.apply(h, torch.float32), c)
    h = _call(enc.conv5, h)
    feat = Fn.AvgPoolFn.apply(h, h._dp_c).view(x.size(0), -1)
    return feat, model.linear(feat)


def run(kind: str, B: int, alpha: float, out):
    torch.manual_seed(42)
    model = R2Plus1DClassifier((3, 21, 128, 128), 2, LS, False, alpha)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    x, y = (port.synthetic_clips if kind == "noise" else port.structured_clips)(B, 21, 128, 128)
    y[0], y[1] = 0, 1
    w = dp_b200.rw_class_weights([300, 17000])
    xd, yd = x.to(DEV), y.to(DEV)

    # ---- fp64 oracle on the device, with taps and gradients ----
    st64 = {k: v.to(DEV) for k, v in port.clone_state(state, dtype=torch.float64).items()}
    for v in st64.values():
        if v.is_floating_point() and v.dtype == torch.float64 and not v.requires_grad and v.is_leaf:
            pass
    st64 = {k: (v.detach().requires_grad_(True) if k.endswith(("weight", "bias")) else v.detach()) for k, v in st64.items()}
    taps = {}
    logits64 = port.classifier_forward(st64, xd.double(), LS, alpha, True, taps=taps)
    loss64 = port.focal_loss(logits64, yd, w.to(DEV).double(), 2.0)
    loss64.backward()
    g64 = {k: v.grad for k, v in st64.items() if v.requires_grad and v.grad is not None}
    taps = {k: v.detach() for k, v in taps.items()}
    with torch.no_grad():
        feat64 = torch.nn.functional.adaptive_avg_pool3d(taps["res2plus1d.conv5.block1"], 1).view(B, -1)

    # ---- bf16-storage oracle (fp32 math), same device ----
    st32 = {k: v.to(DEV) for k, v in port.clone_state(state).items()}
    st32 = {k: (v.detach().requires_grad_(True) if k.endswith(("weight", "bias")) else v.detach()) for k, v in st32.items()}
    taps_e = {}
    with torch.no_grad():
        logits_e = port.classifier_forward(st32, xd, LS, alpha, True, taps=taps_e, storage="bf16")
        feat_e = torch.nn.functional.adaptive_avg_pool3d(taps_e["res2plus1d.conv5.block1"], 1).view(B, -1)

    # ---- CUDA path, bf16, module by module with hooks: accumulated error per activation ----
    m = R2Plus1DClassifier((3, 21, 128, 128), 2, LS, False, alpha)
    m.load_state_dict(state)
    m = m.to(DEV).train()
    acc_err = {}
    handles = []
    from dp_b200.R2Plus1D import Conv3dBlock, SpatioTemporalResBlock
    for name, mod in m.named_modules():
        if isinstance(mod, (Conv3dBlock, SpatioTemporalResBlock)):
            def hook(_m, _i, o, name=name):
                if name in taps:
                    acc_err[name] = rel(o, taps[name])
            handles.append(mod.register_forward_hook(hook))
    with dp_b200.compute_mode("bf16"), torch.no_grad():
        m(xd)
    for h in handles:
        h.remove()

    # ---- local error: each Conv3dBlock alone on the oracle's input ----
    stem, blocks = port.encoder_plan(LS, alpha)
    chain = []          # (layer name, name of the tap that is its input, or None for the clip)
    prev = None
    for lay in stem:
        chain.append((lay[0], prev))
        prev = lay[0]
    for b in blocks:
        block_in = prev
        p = block_in
        for lay in b["conv1"] + b["conv2"]:
            chain.append((lay[0], p))
            p = lay[0]
        if b["shortcut"] is not None:
            p = block_in
            for lay in b["shortcut"]:
                chain.append((lay[0], p))
                p = lay[0]
        prev = b["prefix"]
    mods = dict(m.named_modules())
    loc_err = {}
    with dp_b200.compute_mode("bf16"), torch.no_grad():
        for name, src in chain:
            xin = xd if src is None else taps[src].float()
            o = mods[name](xin)
            loc_err[name] = rel(o, taps[name])
            del o

    # ---- fused product path: features, logits, loss, gradients ----
    m2 = R2Plus1DClassifier((3, 21, 128, 128), 2, LS, False, alpha)
    m2.load_state_dict(state)
    m2 = m2.to(DEV).train()
    lf = FocalLoss(weight=w.to(DEV), gamma=2.0)
    with dp_b200.compute_mode("bf16"):
        feat = m2.res2plus1d(xd)
        logits = m2.linear(feat)
        loss = lf(logits, yd)
        loss.backward()
    st_head = {k: v.detach() for k, v in st64.items()}
    with torch.no_grad():
        logits_head64 = port.head_forward({k: v.clone() for k, v in st_head.items()}, feat.detach().double(), alpha, True)
    gerr = {}
    for n, p_ in m2.named_parameters():
        if n.endswith("conv.weight") and n in g64:
            gerr[n] = rel(p_.grad, g64[n])
    # ---- what-if: fp32 tail ----
    m3 = R2Plus1DClassifier((3, 21, 128, 128), 2, LS, False, alpha)
    m3.load_state_dict(state)
    m3 = m3.to(DEV).train()
    with dp_b200.compute_mode("bf16"), torch.no_grad():
        feat_t, logits_t = forward_fp32_tail(m3, xd)

    def q(vals, f):
        s = sorted(vals)
        return s[min(len(s) - 1, int(f * len(s)))]

    # head conditioning: spread of the features over the batch relative to their size
    spread = (feat64.std(0) / feat64.abs().mean(0).clamp_min(1e-30)).median().item()
    out.write(f"\n## {kind} clips, B = {B}, alpha = {alpha}\n\n")
    out.write(f"| quantity | CUDA bf16 vs fp64 | bf16-storage oracle vs fp64 |\n|---|---|---|\n")
    out.write(f"| pooled features (B,128), rel-L2 | {rel(feat, feat64):.2e} | {rel(feat_e, feat64):.2e} |\n")
    lmax = logits64.abs().max().item()
    out.write(f"| logits, max-abs / max-abs (north_star's metric) | {((logits.double() - logits64).abs().max() / lmax).item():.2e} | "
              f"{((logits_e.double() - logits64).abs().max() / lmax).item():.2e} |\n")
    out.write(f"| logits, rel-L2 | {rel(logits, logits64):.2e} | {rel(logits_e, logits64):.2e} |\n")
    out.write(f"| loss, relative | {abs(loss.item() - loss64.item()) / abs(loss64.item()):.2e} | -- |\n")
    out.write(f"| logits of the fp64 HEAD applied to the CUDA features (head amplification alone), max-abs/max-abs | "
              f"{((logits_head64 - logits64).abs().max() / lmax).item():.2e} | -- |\n")
    out.write(f"| what-if conv5 + pool + head stored in fp32: features rel-L2 / logits max-abs | {rel(feat_t, feat64):.2e} / "
              f"{((logits_t.double() - logits64).abs().max() / lmax).item():.2e} | -- |\n")
    out.write(f"| batch spread of the pooled features (median over channels of std_batch / mean abs) | {spread:.2e} | |\n")
    ge = list(gerr.values())
    out.write(f"| conv weight gradients vs fp64, rel-L2 median / p90 / max | {q(ge, .5):.2e} / {q(ge, .9):.2e} / {max(ge):.2e} | -- |\n")
    out.write("\n| activation | accumulated rel-L2 error | local rel-L2 error (layer alone, oracle input) | weight-gradient rel-L2 |\n|---|---|---|---|\n")
    for name in list(dict.fromkeys([c[0] for c in chain] + [b["prefix"] for b in blocks])):
        a = acc_err.get(name)
        lo = loc_err.get(name)
        ge_ = gerr.get(name + ".conv.weight")
        out.write(f"| {name.replace('res2plus1d.', '')} | {'' if a is None else f'{a:.2e}'} | {'' if lo is None else f'{lo:.2e}'} | "
                  f"{'' if ge_ is None else f'{ge_:.2e}'} |\n")
    out.flush()
    print(f"[{kind} B={B}] features {rel(feat, feat64):.2e} logits {((logits.double() - logits64).abs().max() / lmax).item():.2e} "
          f"head-only {((logits_head64 - logits64).abs().max() / lmax).item():.2e} fp32-tail {((logits_t.double() - logits64).abs().max() / lmax).item():.2e}")
    del taps, taps_e, st64, st32
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--alpha", type=float, default=1.0)
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    path = args.out or os.path.join(ROOT, "gpurun_out", f"{args.tag}_error_attribution.md")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as out:
        out.write("# Per-layer error attribution of the bf16 product mode (scripts/error_attribution.py)\n\n"
                  "Oracle: oracle/r2plus1d_port.py in float64 on the GPU (torch eager, TF32 off).  'accumulated' = the activation of the\n"
                  "CUDA path (module-by-module, hooks) against the oracle's; 'local' = the same Conv3dBlock alone, fed the oracle's input:\n"
                  "what ONE layer adds (two bf16 roundings: raw conv output and activation, 2^-9 relative each).\n")
        for kind in ("noise", "structured"):
            for B in sorted({8, args.batch}):
                run(kind, B, args.alpha, out)
    print("wrote", path)


if __name__ == "__main__":
    main()
