"""Scratch: error of the CUDA path vs the oracle port for noise vs structured clips (DESIGN.md Numerics)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dp_b200
from dp_b200.R2Plus1D import R2Plus1DClassifier
from dp_b200.loss import FocalLoss
from oracle import r2plus1d_port as port

torch.set_num_threads(os.cpu_count())
layer_sizes = [1, 2, 2, 1]
w = dp_b200.rw_class_weights([300, 17000])
def rl2(a, b): return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30)).item()
def rmax(a, b): return ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().cpu().abs().max()).item()
for alpha in (1.0, 0.01):
  for kind in ("noise", "structured"):
    for B in (8,):
        torch.manual_seed(42)
        model = R2Plus1DClassifier((3, 21, 128, 128), 2, layer_sizes, False, alpha)
        state = {k: v.clone() for k, v in model.state_dict().items()}
        x, y = (port.synthetic_clips if kind == "noise" else port.structured_clips)(B)
        y[0], y[1] = 0, 1
        refs = {}
        for name, dt, storage in (("f64", torch.float64, "fp32"), ("f32", None, "fp32"), ("bf16emu", None, "bf16")):
            st = port.clone_state(state, dtype=dt)
            xx = x.to(dt) if dt else x
            taps = {}
            feat = port.encoder_forward(st, xx, layer_sizes, alpha, True, taps, storage)
            logits = port.head_forward(st, feat, alpha, True)
            loss = port.focal_loss(logits, y, w.to(logits.dtype), 2.0)
            loss.backward()
            refs[name] = (feat.detach(), logits.detach(), loss.detach(), {k: v.grad for k, v in st.items() if v.requires_grad})
        f64 = refs["f64"]
        gmax = max(g.norm().item() for g in f64[3].values())
        def report(tag, feat, logits, loss, grads):
            ge = sorted((rl2(grads[n], f64[3][n]), n) for n in grads if f64[3][n].norm().item() > 1e-5 * gmax)
            p_ref = torch.softmax(f64[1], 1)[:, 0]; p = torch.softmax(logits.double().cpu(), 1)[:, 0]
            print(f"a={alpha} {kind} B={B} {tag:10s}: feat {rmax(feat, f64[0]):.2e} logits {rmax(logits, f64[1]):.2e} loss {abs(loss.item()-f64[2].item())/abs(f64[2].item()):.2e} "
                  f"dP {(p-p_ref).abs().max().item():.2e} | grad relL2 median {ge[len(ge)//2][0]:.2e} p90 {ge[int(len(ge)*0.9)][0]:.2e} worst {ge[-1][0]:.2e} ({ge[-1][1]})", flush=True)
        report("ref-f32", *refs["f32"])
        report("ref-bf16emu", *refs["bf16emu"])
        for mode, impl in (("fp32", "simt"), ("bf16", "auto")):
            m = R2Plus1DClassifier((3, 21, 128, 128), 2, layer_sizes, False, alpha)
            m.load_state_dict(state)
            m = m.cuda().train()
            lf = FocalLoss(weight=w.cuda(), gamma=2.0)
            with dp_b200.compute_mode(mode, impl):
                feat = m.res2plus1d(x.cuda())
                logits = m.linear(feat)
                loss = lf(logits, y.cuda())
                loss.backward()
            torch.cuda.synchronize()
            report(f"ours-{mode}", feat.detach(), logits.detach(), loss.detach(), {n: p.grad for n, p in m.named_parameters()})
