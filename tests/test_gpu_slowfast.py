"""GPU: the drop-in SlowFast (config 3) against golden vectors produced by the UNMODIFIED reference
(tests/golden/slowfast_step.npz, oracle/make_golden.py::slowfast_golden)."""
import os

import numpy as np
import pytest
import torch

import dp_b200
from dp_b200.loss import FocalLoss
from dp_b200.slowfast import Bottleneck3D, SlowFast

pytestmark = pytest.mark.gpu
DEV = "cuda"


def summarise(t):
    f = t.detach().double().reshape(-1).cpu()
    head = torch.zeros(4, dtype=torch.float64)
    head[:min(4, f.numel())] = f[:4]
    return np.concatenate([[f.sum().item(), f.abs().sum().item()], head.numpy()])


def _model_and_data(gold):
    T, H, W, B = 20, 64, 64, 4
    torch.manual_seed(42)
    m = SlowFast((3, T, H, W), Bottleneck3D, [1, 2, 2, 1], 4, 1, 2, 1.0)
    # the reference constructor leaves rounding-noise driven BatchNorm buffers: take them from the fixture
    sd = m.state_dict()
    off = 0
    for k in gold["init_bn_keys"]:
        n = sd[str(k)].numel()
        sd[str(k)].copy_(torch.from_numpy(gold["init_bn_values"][off:off + n]).view_as(sd[str(k)]))
        off += n
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 256, (B, 3, T, H, W), generator=g).float() - torch.tensor([90.0, 98.0, 102.0]).view(1, 3, 1, 1, 1)
    y = torch.from_numpy(gold["y"])
    return m, x, y


def test_slowfast_fp32_matches_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "slowfast_step.npz"))
    m, x, y = _model_and_data(gold)
    m = m.to(DEV).train()
    lf = FocalLoss(weight=torch.tensor([0.98, 0.02], device=DEV), gamma=2.0)
    with dp_b200.compute_mode("fp32"):
        logits = m(x.to(DEV))
        loss = lf(logits, y.to(DEV))
        loss.backward()
    ref = torch.from_numpy(gold["logits"])
    e_logit = ((logits.detach().cpu() - ref).abs().max() / ref.abs().max()).item()
    e_loss = abs(loss.item() - float(gold["loss"])) / abs(float(gold["loss"]))
    print(f"[slowfast fp32] logits rel {e_logit:.2e} loss rel {e_loss:.2e}")
    assert e_logit < 1e-4 and e_loss < 1e-4
    params = dict(m.named_parameters())
    gn = gold["grad_norm"]
    worst = 0.0
    for i, n in enumerate(gold["grad_names"]):
        g = params[str(n)].grad
        if gn[i] < 1e-4 * gn.max():          # biases in front of a BatchNorm: zero up to rounding in the reference
            assert g is None or g.double().norm().item() < 1e-3 * gn.max(), n
            continue
        assert g is not None, n
        err = abs(g.double().norm().item() - gn[i]) / gn[i]
        worst = max(worst, err)
        assert err < 2e-2, (n, err)
    print(f"[slowfast fp32] worst gradient-norm deviation {worst:.2e}")
    sd = m.state_dict()
    for k, want in zip(gold["init_bn_keys"], gold["bn_summary"]):
        got = summarise(sd[str(k)])
        assert np.allclose(got, want, rtol=2e-3, atol=2e-4), (k, got, want)
    m.eval()
    with dp_b200.compute_mode("fp32"), torch.no_grad():
        ev = m(x.to(DEV)).cpu()
    ev_ref = torch.from_numpy(gold["eval_logits"])
    assert ((ev - ev_ref).abs().max() / ev_ref.abs().max()).item() < 1e-3


def test_slowfast_bf16_runs_on_tensor_cores_and_trains(golden_dir):
    gold = np.load(os.path.join(golden_dir, "slowfast_step.npz"))
    m, x, y = _model_and_data(gold)
    m = m.to(DEV).train()
    lf = FocalLoss(weight=torch.tensor([0.98, 0.02], device=DEV), gamma=2.0)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    n0 = dp_b200._lib.load().dp_launch_count()
    with dp_b200.compute_mode("bf16", "auto"):
        logits = m(x.to(DEV))
        loss = lf(logits, y.to(DEV))
        loss.backward()
        opt.step()
    assert dp_b200._lib.load().dp_launch_count() - n0 > 300      # 74 convs x (fwd, BN, dgrad, wgrad, ...)
    ref = torch.from_numpy(gold["logits"])
    e_logit = ((logits.detach().cpu() - ref).abs().max() / ref.abs().max()).item()
    e_loss = abs(loss.item() - float(gold["loss"])) / abs(float(gold["loss"]))
    print(f"[slowfast bf16] logits rel {e_logit:.2e} loss rel {e_loss:.2e} vs the fp32 reference")
    assert e_logit < 0.25 and e_loss < 0.1
    moved = sum(int(not torch.equal(before[n], p.detach())) for n, p in m.named_parameters())
    assert moved > 0.9 * len(before)
    assert all(torch.isfinite(p).all() for p in m.parameters())
    assert m.encode(x[:2].to(DEV)).shape == (2, 640)
