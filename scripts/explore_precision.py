"""Scratch: bf16 error of the CUDA path vs the fp32 oracle port as a function of batch size."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dp_b200
from dp_b200.R2Plus1D import R2Plus1DClassifier
from dp_b200.loss import FocalLoss
from oracle import r2plus1d_port as port

torch.set_num_threads(os.cpu_count())
layer_sizes, alpha = [1, 2, 2, 1], float(os.environ.get("ALPHA", "1.0"))
w = dp_b200.rw_class_weights([300, 17000])
for B in [int(b) for b in os.environ.get("BS", "4,8,16,32").split(",")]:
    torch.manual_seed(42)
    model = R2Plus1DClassifier((3, 21, 128, 128), 2, layer_sizes, False, alpha)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    x, y = port.synthetic_clips(B)
    y[0], y[1] = 0, 1
    t0 = time.time()
    st = port.clone_state(state)
    taps = {}
    feat_ref = port.encoder_forward(st, x, layer_sizes, alpha, True, taps)
    logits_ref = port.head_forward(st, feat_ref, alpha, True)
    loss_ref = port.focal_loss(logits_ref, y, w, 2.0)
    loss_ref.backward()
    t_ref = time.time() - t0
    for mode, impl in (("fp32", "simt"), ("bf16", "simt"), ("bf16", "auto")):
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
        ef = ((feat.detach().cpu() - feat_ref.detach()).abs().max() / feat_ref.detach().abs().max()).item()
        el = ((logits.detach().cpu() - logits_ref.detach()).abs().max() / logits_ref.detach().abs().max()).item()
        eloss = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
        gerrs = []
        for n, p in m.named_parameters():
            rg = st[n].grad
            if rg.norm() > 1e-6:
                gerrs.append((((p.grad.cpu() - rg).norm() / rg.norm()).item(), n))
        gerrs.sort(reverse=True)
        print(f"B={B} {mode}/{impl}: feat {ef:.2e} logits {el:.2e} loss {eloss:.2e} | grad relL2 worst {gerrs[0][0]:.2e} ({gerrs[0][1]}) median {gerrs[len(gerrs)//2][0]:.2e} | ref {t_ref:.1f}s", flush=True)
