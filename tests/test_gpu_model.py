"""End-to-end parity of the CUDA path (through the drop-in nn.Modules, i.e. through the C ABI) against
(1) the golden vectors produced by the UNMODIFIED reference (tests/golden/small_train_step.npz) and
(2) the oracle port (oracle/r2plus1d_port.py) on the same seeded state and inputs.

Stated tolerances:
  * fp32 validation mode: logits and loss within 1e-4 relative of the reference (north_star); gradients per
    tensor against an fp64 run of the oracle: our error <= max(1e-4, 5 x the fp32 oracle's own error)
    (the train-mode-BN network is ill-conditioned on noise clips: fp32 summation order alone moves
    gradients by ~1e-2 relative L2, see DESIGN.md "Numerics").
  * bf16 product mode: the checker is the oracle with bf16 STORAGE emulation (same fp32 algorithm, tensors
    rounded to bf16 where the CUDA path stores them): logits/loss within 1e-2 relative (north_star's bf16
    bound), gradients within BF16_GRAD_TOL relative L2.  Against the pure-fp32 oracle bf16 storage itself
    costs 2e-2 (features) to ~1e-1 (logits) on these inputs -- measured identically for the emulated
    reference -- so that comparison is asserted only against the envelope BF16_VS_FP32.
  * >= 99.9 % agreement on thresholded disruption labels."""
import os
import sys

import numpy as np
import pytest
import torch

import dp_b200
from dp_b200 import functional as Fn
from dp_b200.R2Plus1D import R2Plus1DClassifier, Conv3dBlock, SpatioTemporalConv
from dp_b200.loss import FocalLoss, LDAMLoss, CELoss

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import r2plus1d_port as port  # noqa: E402  (the checker, never the thing measured)

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_GRAD_TOL = 5e-2
BF16_VS_FP32 = {"logits": 0.3, "loss": 0.2}
CLS = [300, 17000]


def rel_max(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def summarise(t):
    f = t.detach().double().reshape(-1).cpu()
    head = torch.zeros(4, dtype=torch.float64)
    head[:min(4, f.numel())] = f[:4]
    return np.concatenate([[f.sum().item(), f.abs().sum().item()], head.numpy()])


def build(input_size, layer_sizes, alpha, seed=42):
    torch.manual_seed(seed)
    return R2Plus1DClassifier(input_size, 2, layer_sizes, False, alpha)


def make_loss(name, w, s=1.0):
    if name == "focal":
        return FocalLoss(weight=w, gamma=2.0)
    if name == "ldam":
        return LDAMLoss(CLS, max_m=0.5, weight=w, s=s)
    return CELoss(weight=w)


@pytest.mark.parametrize("alpha", [1.0, 0.01])
@pytest.mark.parametrize("loss_name", ["focal", "ldam", "ce"])
def test_fp32_mode_matches_reference_golden(golden_dir, alpha, loss_name):
    """CUDA fp32 validation mode vs numbers the reference itself produced (make_golden.py, case B)."""
    gold = np.load(os.path.join(golden_dir, "small_train_step.npz"))
    B, T, H, W = 4, 5, 32, 32
    x, y = port.synthetic_clips(B, T, H, W)
    y[0], y[1] = 0, 1
    assert np.allclose(summarise(x), gold["x_summary"])
    model = build((3, T, H, W), [1, 1, 1, 1], alpha).to(DEV).train()
    w = dp_b200.rw_class_weights(CLS).to(DEV)
    lf = make_loss(loss_name, w)
    tag = f"a{alpha}_{loss_name}"
    with dp_b200.compute_mode("fp32"):
        logits = model(x.to(DEV))
        loss = lf(logits, y.to(DEV))
        loss.backward()
        assert rel_max(logits, torch.from_numpy(gold[tag + "_logits"])) < 1e-4
        assert abs(loss.item() - float(gold[tag + "_loss"])) / abs(float(gold[tag + "_loss"])) < 1e-4
        names = list(gold[tag + "_grad_names"])
        params = dict(model.named_parameters())
        gn = gold[tag + "_grad_norm"]
        gs = gold[tag + "_grad_summary"]
        for i, n in enumerate(names):
            g = params[n].grad
            assert g is not None, n
            norm = g.double().norm().item()
            assert abs(norm - gn[i]) <= 2e-3 * max(gn.max() * 1e-3, gn[i]), (n, norm, gn[i])
            got = summarise(g)
            if gn[i] > 1e-4 * gn.max():   # skip tensors whose true gradient is zero (bias in front of a BN)
                scale = max(gs[i][1] / g.numel(), np.abs(gs[i][2:]).max())   # mean |grad| vs sampled values
                assert np.all(np.abs(got[2:] - gs[i][2:]) <= 3e-2 * scale), (n, got, gs[i])
        sd = model.state_dict()
        for k, ref in zip(gold[tag + "_bn_keys"], gold[tag + "_bn_summary"]):
            got = summarise(sd[str(k)])
            assert np.allclose(got, ref, rtol=2e-4, atol=1e-6), k
        model.eval()
        with torch.no_grad():
            ev = model(x.to(DEV))
        assert rel_max(ev, torch.from_numpy(gold[tag + "_eval_logits"])) < 1e-4


def _port_step(state, x, y, layer_sizes, alpha, loss_name, w, s=1.0, storage="fp32", dtype=None):
    st = port.clone_state(state, dtype=dtype)
    margins = port.ldam_margins(CLS, 0.5)
    if dtype is not None:
        x, w, margins = x.to(dtype), w.to(dtype), margins.to(dtype)
    return port.train_step(st, x, y, layer_sizes, alpha, loss=loss_name, weight=w, margins=margins, s=s,
                           storage=storage) + (st,)


def _cuda_step(state, x, y, layer_sizes, alpha, loss_name, w, mode, impl="auto"):
    model = build((3, x.shape[2], x.shape[3], x.shape[4]), layer_sizes, alpha)
    model.load_state_dict(state)
    model = model.to(DEV).train()
    lf = make_loss(loss_name, w.to(DEV))
    with dp_b200.compute_mode(mode, impl):
        logits = model(x.to(DEV))
        loss = lf(logits, y.to(DEV))
        loss.backward()
    torch.cuda.synchronize()
    return model, logits.detach(), loss.detach()


def test_full_size_fp32_mode_vs_oracle():
    """Benchmark model ([1,2,2,1], clips (3,21,128,128)), fp32 validation mode: fwd + Focal + bwd."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    B, layer_sizes, alpha = 4, [1, 2, 2, 1], 1.0
    x, y = port.synthetic_clips(B)
    y[0], y[1] = 0, 1
    state = {k: v.clone() for k, v in build((3, 21, 128, 128), layer_sizes, alpha).state_dict().items()}
    w = dp_b200.rw_class_weights(CLS)
    ref_logits, ref_loss, ref_grads, ref_state = _port_step(state, x, y, layer_sizes, alpha, "focal", w)
    _, _, g64, _ = _port_step(state, x, y, layer_sizes, alpha, "focal", w, dtype=torch.float64)
    model, logits, loss = _cuda_step(state, x, y, layer_sizes, alpha, "focal", w, "fp32")
    e_logit, e_loss = rel_max(logits, ref_logits), abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
    print(f"[fp32] logits rel {e_logit:.3e} loss rel {e_loss:.3e}")
    assert e_logit < 1e-4 and e_loss < 1e-4
    worst = ("", 0.0, 0.0)
    gmax = max(g.double().norm().item() for g in g64.values())
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        if g64[n].norm().item() < 1e-6 * gmax:
            assert p.grad.double().norm().item() < 1e-4 * gmax, n    # structurally-zero gradient stays ~0
            continue
        e_ours, e_ref = rel_l2(p.grad, g64[n]), rel_l2(ref_grads[n], g64[n])
        assert e_ours <= max(1e-4, 5.0 * e_ref), (n, e_ours, e_ref)
        if e_ours > worst[1]:
            worst = (n, e_ours, e_ref)
    print(f"[fp32] worst gradient error vs fp64 oracle {worst[1]:.3e} (fp32 oracle itself {worst[2]:.3e}) at {worst[0]}")
    sd = model.state_dict()
    for k in sd:
        if k.endswith("running_var") or k.endswith("running_mean"):
            assert rel_max(sd[k], ref_state[k]) < 1e-4, k
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(ref_state[k]), k


@pytest.mark.parametrize("loss_name,alpha,impl", [
    ("focal", 1.0, "auto"),
    ("ldam", 0.01, "auto"),
    ("focal", 0.0, "auto"),
    ("ce", 1.0, "simt"),
])
def test_full_size_bf16_mode_vs_oracle(loss_name, alpha, impl):
    """bf16 product path (tcgen05 kernels where covered) against the bf16-storage oracle and, as an envelope,
    against the fp32 oracle."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    B, layer_sizes = 4, [1, 2, 2, 1]
    x, y = port.synthetic_clips(B)
    y[0], y[1] = 0, 1
    state = {k: v.clone() for k, v in build((3, 21, 128, 128), layer_sizes, alpha).state_dict().items()}
    w = dp_b200.rw_class_weights(CLS)
    f32_logits, f32_loss, _, _ = _port_step(state, x, y, layer_sizes, alpha, loss_name, w)
    ref_logits, ref_loss, ref_grads, ref_state = _port_step(state, x, y, layer_sizes, alpha, loss_name, w,
                                                             storage="bf16")
    model, logits, loss = _cuda_step(state, x, y, layer_sizes, alpha, loss_name, w, "bf16", impl)
    e_logit, e_loss = rel_max(logits, ref_logits), abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
    v_logit, v_loss = rel_max(logits, f32_logits), abs(loss.item() - f32_loss.item()) / abs(f32_loss.item())
    o_logit = rel_max(ref_logits, f32_logits)
    print(f"[bf16 {loss_name} a={alpha} {impl}] vs bf16-storage oracle: logits {e_logit:.3e} loss {e_loss:.3e} | "
          f"vs fp32 oracle: logits {v_logit:.3e} loss {v_loss:.3e} (bf16-storage oracle itself: {o_logit:.3e})")
    assert e_logit < 1e-2 and e_loss < 1e-2
    assert v_logit < BF16_VS_FP32["logits"] and v_loss < BF16_VS_FP32["loss"]
    worst = ("", 0.0)
    gmax = max(g.double().norm().item() for g in ref_grads.values())
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        rg = ref_grads[n]
        if rg.double().norm().item() < 1e-5 * gmax:
            continue
        e = rel_l2(p.grad, rg)
        assert e < BF16_GRAD_TOL, (n, e)
        if e > worst[1]:
            worst = (n, e)
    print(f"[bf16] worst gradient rel-L2 error vs bf16-storage oracle {worst[1]:.3e} at {worst[0]}")
    sd = model.state_dict()
    for k in sd:
        if k.endswith("running_var") or k.endswith("running_mean"):
            assert rel_max(sd[k], ref_state[k]) < 5e-3, k


def test_thresholded_label_agreement():
    """not(softmax[:,0] > thr) labels (reference src/evaluate.py:56-57), eval mode, 64 clips."""
    layer_sizes, alpha = [1, 2, 2, 1], 1.0
    model = build((3, 21, 128, 128), layer_sizes, alpha)
    # give BN non-trivial running statistics: one train step on the oracle
    state = {k: v.clone() for k, v in model.state_dict().items()}
    st = port.clone_state(state, requires_grad=False)
    xs, _ = port.synthetic_clips(4, seed=99)
    with torch.no_grad():
        port.classifier_forward(st, xs, layer_sizes, alpha, training=True)
    model.load_state_dict({k: v.detach() for k, v in st.items()})
    model = model.to(DEV).eval()
    agree = {"fp32": 0, "bf16": 0}
    confident = {"fp32": [0, 0], "bf16": [0, 0]}
    total = 0
    for chunk in range(4):
        x, _ = port.synthetic_clips(16, seed=500 + chunk)
        with torch.no_grad():
            ref = port.classifier_forward(st, x, layer_sizes, alpha, training=False)
        ref_lab = ~(torch.softmax(ref, 1)[:, 0] > 0.5)
        margin = (ref[:, 0] - ref[:, 1]).abs()
        for mode in ("fp32", "bf16"):
            with dp_b200.compute_mode(mode), torch.no_grad():
                out = model(x.to(DEV)).cpu()
            lab = ~(torch.softmax(out, 1)[:, 0] > 0.5)
            agree[mode] += int((lab == ref_lab).sum())
            sure = margin > 0.02 * ref.abs().max()
            confident[mode][0] += int(((lab == ref_lab) & sure).sum())
            confident[mode][1] += int(sure.sum())
        total += 16
    print("label agreement", {m: agree[m] / total for m in agree}, "confident", confident)
    assert agree["fp32"] / total >= 0.999
    assert confident["bf16"][0] == confident["bf16"][1]   # 100 % on clips not sitting on the threshold
    assert agree["bf16"] / total >= 0.95


def test_state_dict_and_module_surface():
    model = build((3, 21, 128, 128), [1, 2, 2, 1], 1.0)
    sd = model.state_dict()
    assert len(sd) == 201
    assert sd["res2plus1d.conv1.spatio_conv.conv.weight"].shape == (45, 3, 1, 7, 7)
    assert sd["res2plus1d.conv3.block1.downsample_conv.spatio_conv.conv.weight"].shape == (21, 32, 1, 1, 1)
    model2 = build((3, 21, 128, 128), [1, 2, 2, 1], 1.0, seed=7)
    model2.load_state_dict(sd)
    model.to(DEV)
    x, _ = port.synthetic_clips(2)
    feat = model.encode(x.to(DEV))
    assert feat.shape == (2, 128) and not feat.requires_grad
    with pytest.raises(dp_b200._lib.DpError):
        model.cpu()(x)     # no CPU path: fails loudly
    model.to(DEV)


def test_hooks_and_standalone_modules_see_ncdhw():
    """GradCAM-style hooks on res2plus1d.conv5 (reference visualize_cam.py:75-95) get NCDHW fp32 tensors and
    the hooked run gives the same logits/gradients as the fused run."""
    torch.manual_seed(0)
    model = build((3, 9, 64, 64), [1, 1, 1, 1], 0.01).to(DEV).train()
    x, y = port.synthetic_clips(2, 9, 64, 64)
    x, y = x.to(DEV), y.to(DEV)
    lf = CELoss(weight=torch.ones(2, device=DEV))
    seen = {}

    def fwd_hook(mod, inp, out):
        seen["act"] = out.detach()

    def bwd_hook(mod, gin, gout):
        seen["grad"] = gout[0].detach()

    with dp_b200.compute_mode("fp32"):
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        out_a = model(x)
        lf(out_a, y).backward()
        ga = {n: p.grad.clone() for n, p in model.named_parameters()}
        model.zero_grad()
        model.load_state_dict(sd0)
        h1 = model.res2plus1d.conv5.register_forward_hook(fwd_hook)
        h2 = model.res2plus1d.conv5.register_full_backward_hook(bwd_hook)
        h3 = model.res2plus1d.conv3.block1.conv1.spatio_conv.register_forward_hook(lambda m, i, o: seen.__setitem__("inner", o.shape))
        out_b = model(x)
        lf(out_b, y).backward()
        h1.remove(); h2.remove(); h3.remove()
    assert seen["act"].shape == (2, 128, 2, 4, 4) and seen["act"].dtype == torch.float32
    assert seen["grad"].shape == (2, 128, 2, 4, 4)
    assert seen["inner"] == (2, 115, 9, 16, 16)
    assert rel_max(out_b, out_a) < 1e-5
    for n, p in model.named_parameters():
        assert rel_max(p.grad, ga[n]) < 1e-4, n
    # stand-alone blocks are drop-ins on NCDHW tensors
    blk = SpatioTemporalConv(8, 16, (3, 3, 3), (1, 1, 1), 1, (1, 1, 1), alpha=0.2).to(DEV)
    xin = torch.randn(2, 8, 4, 10, 10, device=DEV)
    with dp_b200.compute_mode("fp32"):
        out = blk(xin)
    assert out.shape == (2, 16, 4, 10, 10) and out.dtype == torch.float32


def test_train_loop_smoke_params_move_no_nan():
    """Intent of the reference's torcheck smoke test (test/test_model.py:49-162): after optimiser steps every
    parameter changed, nothing is NaN/Inf, logits are not confined to (0,1)."""
    from dp_b200.optim import FusedClipAdamW
    model = build((3, 21, 128, 128), [1, 2, 2, 1], 1.0).to(DEV).train()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    opt = FusedClipAdamW(model.parameters(), lr=2e-4, max_norm=1.0)
    lf = FocalLoss(weight=dp_b200.rw_class_weights(CLS).to(DEV), gamma=2.0)
    losses = []
    for step in range(3):
        x, y = port.synthetic_clips(8, seed=100 + step)
        y[0], y[1] = 0, 1
        opt.zero_grad()
        out = model(x.to(DEV))
        loss = lf(out, y.to(DEV))
        assert torch.isfinite(loss)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    for n, p in model.named_parameters():
        assert torch.isfinite(p).all(), n
        assert not torch.equal(p.detach(), before[n]), f"{n} did not change"
    assert (out.detach().abs() > 1).any() or (out.detach() < 0).any()
    print("losses", losses)
