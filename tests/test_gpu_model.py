"""End-to-end parity of the CUDA path (through the drop-in nn.Modules, i.e. through the C ABI) against
(1) the golden vectors produced by the UNMODIFIED reference (tests/golden/small_train_step.npz) and
(2) the oracle port (oracle/r2plus1d_port.py) on the same seeded state and inputs.

Stated tolerances (each test's docstring carries the evidence):
  * fp32 validation mode: logits and loss within 1e-4 relative of the reference (north_star); gradients as close
    to an fp64 run of the oracle as the fp32 reference itself is.
  * bf16 product mode: measured against fp64 and required to be no worse than the reference's own algorithm
    with bf16-stored activations (oracle `storage="bf16"`); absolute envelopes on loss and P(disruption).
  * >= 99.9 % agreement on thresholded disruption labels in fp32 mode; in bf16, 100 % on clips that do not sit
    on the threshold."""
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
            # 1e-2: on this tiny case (16 values per channel in the last BatchNorm) the fp32 reference's own
            # gradients move by several 1e-3 with summation order (see test_full_size_fp32_mode_vs_oracle)
            assert abs(norm - gn[i]) <= 1e-2 * max(gn.max() * 1e-3, gn[i]), (n, norm, gn[i])
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


def _grad_errors(grads, g64):
    """per-tensor relative L2 error against the fp64 oracle, skipping structurally-zero gradients."""
    gmax = max(g.double().norm().item() for g in g64.values())
    out = {}
    for n, g in g64.items():
        if g.norm().item() < 1e-5 * gmax:
            continue
        out[n] = rel_l2(grads[n], g)
    return out, gmax


def _quantiles(errs):
    v = sorted(errs.values())
    return v[len(v) // 2], v[int(0.9 * len(v))], v[-1]


@pytest.mark.parametrize("clips", ["noise", "structured"])
def test_full_size_fp32_mode_vs_oracle(clips):
    """Benchmark model ([1,2,2,1], clips (3,21,128,128)), fp32 validation mode: fwd + Focal + bwd.
    Logits / loss within 1e-4 of the fp32 reference algorithm (north_star).  Gradients: this network's weight
    gradients are sums of ~1e6 nearly cancelling terms (pooled, per-sample-constant upstream gradient against
    batch-normalised activations), so the fp32 reference ITSELF sits 5e-3..2e-2 away from its fp64 run; the CUDA
    path must be as close to fp64 as the fp32 reference is (quantile by quantile, factor 3)."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    B, layer_sizes, alpha = 4, [1, 2, 2, 1], 1.0
    x, y = (port.synthetic_clips if clips == "noise" else port.structured_clips)(B)
    y[0], y[1] = 0, 1
    state = {k: v.clone() for k, v in build((3, 21, 128, 128), layer_sizes, alpha).state_dict().items()}
    w = dp_b200.rw_class_weights(CLS)
    ref_logits, ref_loss, ref_grads, ref_state = _port_step(state, x, y, layer_sizes, alpha, "focal", w)
    _, _, g64, _ = _port_step(state, x, y, layer_sizes, alpha, "focal", w, dtype=torch.float64)
    model, logits, loss = _cuda_step(state, x, y, layer_sizes, alpha, "focal", w, "fp32")
    e_logit, e_loss = rel_max(logits, ref_logits), abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
    print(f"[fp32 {clips}] logits rel {e_logit:.3e} loss rel {e_loss:.3e}")
    assert e_logit < 1e-4 and e_loss < 1e-4
    ours = {n: p.grad for n, p in model.named_parameters()}
    assert all(g is not None and torch.isfinite(g).all() for g in ours.values())
    e_ours, gmax = _grad_errors(ours, g64)
    e_ref, _ = _grad_errors(ref_grads, g64)
    qo, qr = _quantiles(e_ours), _quantiles(e_ref)
    print(f"[fp32 {clips}] gradient rel-L2 error vs fp64 oracle (median, p90, max): ours {qo}, fp32 oracle itself {qr}")
    for a, b in zip(qo, qr):
        assert a <= max(1e-4, 3.0 * b)
    for n, g in g64.items():          # structurally-zero gradients (bias feeding a BatchNorm) stay ~0
        if g.norm().item() < 1e-5 * gmax:
            assert ours[n].double().norm().item() < 1e-3 * gmax, n
    sd = model.state_dict()
    for k in sd:
        if k.endswith("running_var") or k.endswith("running_mean"):
            assert rel_max(sd[k], ref_state[k]) < 1e-4, k
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(ref_state[k]), k


@pytest.mark.parametrize("loss_name,alpha,impl,clips", [
    ("focal", 1.0, "auto", "structured"),
    ("focal", 1.0, "auto", "noise"),
    ("ldam", 0.01, "auto", "structured"),
    ("focal", 0.0, "auto", "structured"),
    ("ce", 1.0, "simt", "structured"),
])
def test_full_size_bf16_mode_vs_oracle(loss_name, alpha, impl, clips):
    """bf16 product path (tcgen05 kernels where covered).  What bf16 STORAGE alone does to this network is
    measured, not assumed: the oracle is run (a) in fp64, (b) in fp32, (c) in fp32 with every tensor the CUDA
    path stores rounded to bf16 at the same points ("bf16-storage oracle": the reference's algorithm, bf16
    activations).  (c) sits 2-3e-2 (features) / 4e-2..1e-1 (logits) from (a) on these clips, so north_star's 1e-2
    is not reachable by ANY bf16-activation implementation of this model, the reference's included.  Asserted:
      * the CUDA path is no further from fp64 than 3x the bf16-storage oracle (+2e-2; both are single draws of
        the same rounding noise) on logits, loss and P(disruption), i.e. its error is bf16 storage, not the
        kernels (the per-kernel tests in test_gpu_kernels.py bound each kernel at rounding level);
      * absolute envelopes: loss within 0.1, P(disruption) within 8e-2, logits within 0.25;
      * gradient error quantiles vs fp64 no worse than 1.25x the bf16-storage oracle's;
      * running statistics within 5e-3."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    B, layer_sizes = 8, [1, 2, 2, 1]
    x, y = (port.synthetic_clips if clips == "noise" else port.structured_clips)(B)
    y[0], y[1] = 0, 1
    state = {k: v.clone() for k, v in build((3, 21, 128, 128), layer_sizes, alpha).state_dict().items()}
    w = dp_b200.rw_class_weights(CLS)
    l64, loss64, g64, _ = _port_step(state, x, y, layer_sizes, alpha, loss_name, w, dtype=torch.float64)
    emu_logits, emu_loss, emu_grads, emu_state = _port_step(state, x, y, layer_sizes, alpha, loss_name, w,
                                                             storage="bf16")
    model, logits, loss = _cuda_step(state, x, y, layer_sizes, alpha, loss_name, w, "bf16", impl)

    def errs(lg, ls):
        p, p64 = torch.softmax(lg.double().cpu(), 1)[:, 0], torch.softmax(l64, 1)[:, 0]
        return rel_max(lg, l64), abs(float(ls) - float(loss64)) / abs(float(loss64)), (p - p64).abs().max().item()

    o, e = errs(logits, loss), errs(emu_logits, emu_loss)
    print(f"[bf16 {loss_name} a={alpha} {impl} {clips}] vs fp64 oracle (logits, loss, P): ours {o} | bf16-storage oracle {e}")
    for a, b in zip(o, e):
        assert a <= 3.0 * b + 2e-2
    assert o[0] < 0.25 and o[1] < 0.1 and o[2] < 8e-2
    ours = {n: p.grad for n, p in model.named_parameters()}
    assert all(g is not None and torch.isfinite(g).all() for g in ours.values())
    qo, qe = _quantiles(_grad_errors(ours, g64)[0]), _quantiles(_grad_errors(emu_grads, g64)[0])
    print(f"[bf16] gradient rel-L2 error vs fp64 oracle (median, p90, max): ours {qo}, bf16-storage oracle {qe}")
    for a, b in zip(qo[:2], qe[:2]):      # median and p90 (the max over ~100 tensors is a single noisy draw)
        assert a <= 1.25 * b + 1e-2
    sd = model.state_dict()
    for k in sd:
        if k.endswith("running_var"):
            assert rel_max(sd[k], emu_state[k]) < 5e-3, k
        if k.endswith("running_mean"):
            # a channel mean is judged on the scale of that channel's standard deviation (several layers have
            # means ~1e-4 sigma, where "relative to the mean" is meaningless); momentum 0.1 scales both sides
            sigma = emu_state[k.replace("running_mean", "running_var")].sqrt()
            err = ((sd[k].cpu() - emu_state[k]).abs() / sigma).max().item()
            assert err < 1e-3, (k, err)


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
