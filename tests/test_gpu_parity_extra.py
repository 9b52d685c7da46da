"""GPU parity cases added in round 2: every loss known-answer case through the CUDA
kernel, BatchNorm statistics at |mean| = 100 sigma, a multi-step fp32 trajectory with the STOCK optimiser (row a12), the
BASELINE-config golden step from the unmodified reference, and strict tcgen05 coverage of all SlowFast convolutions."""
import os
import sys

import numpy as np
import pytest
import torch

import dp_b200
from dp_b200 import _lib
from dp_b200.R2Plus1D import Conv3dBlock, R2Plus1DClassifier
from dp_b200.loss import CELoss, FocalLoss, LDAMLoss

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import r2plus1d_port as port  # noqa: E402  (checker)

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("C_", [2, 5])
def test_loss_kernel_known_answers(golden_dir, C_):
    """dp_loss_fwd_bwd (csrc/loss.cu) against EVERY case of loss_kat.npz (reference src/loss.py:25-34,58-69,80-81 run
    by oracle/make_golden.py): loss values and d loss / d logits."""
    gold = np.load(os.path.join(golden_dir, "loss_kat.npz"))
    target = torch.from_numpy(gold[f"c{C_}_target"]).to(DEV)
    w = torch.from_numpy(gold[f"c{C_}_weight"]).to(DEV)
    counts = gold[f"c{C_}_counts"].tolist()
    cases = {
        "focal_g2": FocalLoss(weight=w, gamma=2.0), "focal_g0": FocalLoss(weight=w, gamma=0.0),
        "focal_g1.5": FocalLoss(weight=w, gamma=1.5),
        "ldam_s30": LDAMLoss(counts, max_m=0.5, weight=w, s=30), "ldam_s1": LDAMLoss(counts, max_m=0.5, weight=w, s=1.0),
        "ldam_s1_now": LDAMLoss(counts, max_m=0.5, weight=None, s=1.0),
        "ce": CELoss(weight=w), "ce_now": CELoss(weight=None),
    }
    for name, lf in cases.items():
        z = torch.from_numpy(gold[f"c{C_}_logits"]).to(DEV).requires_grad_(True)
        val = lf(z, target)
        val.backward()
        ref = float(gold[f"c{C_}_{name}_loss"])
        assert abs(val.item() - ref) <= 5e-6 * max(1.0, abs(ref)), (name, val.item(), ref)
        gref = gold[f"c{C_}_{name}_grad"]
        atol = 2e-5 * max(1.0, float(np.abs(gref).max()))
        np.testing.assert_allclose(z.grad.cpu().numpy(), gref, rtol=5e-5, atol=atol, err_msg=name)
        if name.startswith("ldam"):
            np.testing.assert_array_equal(lf.m_list.numpy(), gold[f"c{C_}_{name}_m"])
    # an upstream gradient other than 1 scales the logit gradient (dp_loss_bwd_scale)
    z = torch.from_numpy(gold[f"c{C_}_logits"]).to(DEV).requires_grad_(True)
    (0.25 * cases["focal_g2"](z, target)).backward()
    np.testing.assert_allclose(z.grad.cpu().numpy(), 0.25 * gold[f"c{C_}_focal_g2_grad"], rtol=5e-5, atol=1e-5)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
@pytest.mark.parametrize("offset", [0.0, 100.0, 1000.0])   # 1000 sigma: reported, outside this network's range (see below)
def test_bn_statistics_large_mean(mode, offset):
    """Train-mode BatchNorm statistics when |mean| >> sigma (SURVEY hard part 2: the un-normalised stem input sits at
    ~ +-100 grey levels).  The conv output is offset + unit noise; mean and (unbiased) variance are read back through the
    running-stat update and compared with float64 statistics of the same bf16-rounded operands."""
    torch.manual_seed(3)
    B, Cin, K, T, H, W = 4, 16, 32, 4, 32, 32
    blk = Conv3dBlock(Cin, K, kernel_size=1, stride=1, padding=0, alpha=1.0).to(DEV).train()
    x = torch.randn(B, Cin, T, H, W, device=DEV)
    x[:, 0] = 1.0                                   # a constant channel carries the offset through a 1x1x1 conv
    with torch.no_grad():
        blk.conv.weight.mul_(0.0).add_(torch.randn_like(blk.conv.weight) / Cin ** 0.5)
        blk.conv.weight[:, 0] = offset
        blk.bn.running_mean.zero_()
        blk.bn.running_var.fill_(1.0)
    with dp_b200.compute_mode(mode), torch.no_grad():
        blk(x)
    torch.cuda.synchronize()
    xr = x.bfloat16().double() if mode == "bf16" else x.double()
    wr = blk.conv.weight.detach().bfloat16().double() if mode == "bf16" else blk.conv.weight.detach().double()
    y = torch.nn.functional.conv3d(xr, wr)
    mean = y.mean(dim=(0, 2, 3, 4))
    var_u = y.var(dim=(0, 2, 3, 4), unbiased=True)
    got_mean = blk.bn.running_mean.double() / 0.1
    got_var = (blk.bn.running_var.double() - 0.9) / 0.1
    e_mean = ((got_mean - mean).abs() / mean.abs().clamp_min(1.0)).max().item()
    e_var = ((got_var - var_u).abs() / var_u).max().item()
    print(f"[{mode} offset {offset}] mean rel err {e_mean:.2e}, variance rel err {e_var:.2e} (sigma ~ 1)")
    assert e_mean < 1e-5
    # E[y^2] - mean^2 from fp32 per-thread partials (fp64 combine) loses ~(mean/sigma)^2 * 2^-24 of relative accuracy:
    # measured 1e-6 at 0, 3e-4..6e-4 at 100 sigma, 2.5e-2 at 1000 sigma.  Every conv of this path has zero-mean (kaiming)
    # weights over BatchNorm-ed or mean-subtracted inputs, so |mean| / sigma stays O(1) (stem input: |mean| <= 0.5 sigma);
    # 100 sigma is the asserted envelope, 1000 sigma is recorded (a shifted sum would cost an extra subtract per
    # accumulator element in the narrow layers' epilogue).
    assert e_var < (2e-3 if offset <= 100.0 else 1e-1), "BatchNorm variance lost accuracy at |mean| >> sigma"


def test_fp32_multi_step_stock_adamw_protocol_matches_port():
    """Row a12: the UNCHANGED caller protocol -- torch.optim.AdamW + clip_grad_norm_ on the drop-in model's .grad
    tensors -- over several optimiser steps against the oracle port driven by the same stock optimiser.

    AdamW's first updates are sign-like (m / sqrt(v) = +-1), so a FREE-RUNNING pair of fp32 trajectories separates at
    the rounding level of the (ill-conditioned) gradients within two steps -- measured here: loss 5e-6, 5e-3, 0.27 apart
    at steps 0, 1, 2 -- which says nothing about the kernels.  The comparison is therefore teacher-forced: before every
    step the CUDA model takes the port's current state; loss, clipped gradient norm, BatchNorm running statistics and the
    weights after the stock optimiser step are compared step by step.  Free-running 300-step trajectories are compared in
    tests/test_gpu_convergence.py (bands, not point-wise)."""
    layer_sizes, alpha, B, T, H, W = [1, 1, 1, 1], 0.01, 4, 9, 64, 64
    n_steps = 5
    w = dp_b200.rw_class_weights([300, 17000])
    torch.manual_seed(42)
    model = R2Plus1DClassifier((3, T, H, W), 2, layer_sizes, False, alpha)
    st = port.clone_state({k: v.clone() for k, v in model.state_dict().items()})
    params = [v for v in st.values() if v.requires_grad]
    opt_ref = torch.optim.AdamW(params, lr=1e-3)
    model = model.to(DEV).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    lf = FocalLoss(weight=w.to(DEV), gamma=2.0)
    worst = {"loss": 0.0, "norm": 0.0, "bn": 0.0, "update": 0.0}
    for i in range(n_steps):
        x, y = port.structured_clips(B, T, H, W, seed=50 + i)
        y[0], y[1] = 0, 1
        model.load_state_dict({k: v.detach().clone() for k, v in st.items()})      # teacher forcing
        before = {k: v.detach().clone() for k, v in st.items() if v.requires_grad}
        opt_ref.zero_grad()
        loss_ref = port.focal_loss(port.classifier_forward(st, x, layer_sizes, alpha, True), y, w, 2.0)
        loss_ref.backward()
        norm_ref = torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt_ref.step()
        with dp_b200.compute_mode("fp32"):
            opt.zero_grad()
            out = model(x.to(DEV))
            loss = lf(out, y.to(DEV))
            assert torch.isfinite(loss)
            loss.backward()
            norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
        worst["loss"] = max(worst["loss"], abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()))
        worst["norm"] = max(worst["norm"], abs(norm.item() - norm_ref.item()) / norm_ref.item())
        sd = model.state_dict()
        for k, v in st.items():
            if k.endswith(("running_mean", "running_var")):
                worst["bn"] = max(worst["bn"], ((sd[k].cpu() - v).abs().max() / v.abs().max().clamp_min(1e-6)).item())
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(v), k
        if i == 0:
            # the very first AdamW update is lr * sign(g): compare the step directions where the gradient is not noise
            agree, total = 0, 0
            for k, v0 in before.items():
                d_ref = (st[k].detach() - v0).reshape(-1)
                d_got = (sd[k].cpu() - v0).reshape(-1)
                big = st[k].grad.reshape(-1).abs() > 1e-2 * st[k].grad.abs().max()
                agree += int((torch.sign(d_ref[big]) == torch.sign(d_got[big])).sum())
                total += int(big.sum())
            worst["update"] = 1.0 - agree / max(1, total)
    print("teacher-forced fp32 steps, worst deviations:", worst)
    assert worst["loss"] < 1e-4 and worst["norm"] < 2e-2 and worst["bn"] < 1e-3
    assert worst["update"] < 5e-3       # fraction of well-conditioned weights whose first update has the other sign


def test_bf16_logit_error_is_the_heads_amplification_of_a_small_feature_error():
    """What the bf16 product mode does at the BASELINE model (see profiles/r2_error_attribution.md for all 32 layers at
    B = 64): the pooled (B,128) features are within 2e-2 relative L2 of a float64 oracle -- one layer adds ~3.5e-3 -- and
    the logit error (north_star's 1e-2 is NOT met: 7e-2 .. 1.2e-1 at random initial weights) is entirely the reference's
    own head, Linear -> BatchNorm1d(batch statistics) -> ELU -> Linear, amplifying that feature error by
    |feature| / std_over_batch ~ 15: the float64 head applied to the CUDA features reproduces the CUDA logits."""
    torch.backends.cudnn.allow_tf32 = False
    B = 16
    torch.manual_seed(42)
    model = R2Plus1DClassifier((3, 21, 128, 128), 2, [1, 2, 2, 1], False, 1.0)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    x, _ = port.structured_clips(B, 21, 128, 128)
    xd = x.to(DEV)
    st64 = {k: v.to(DEV) for k, v in port.clone_state(state, requires_grad=False, dtype=torch.float64).items()}
    with torch.no_grad():
        feat64 = port.encoder_forward(st64, xd.double(), [1, 2, 2, 1], 1.0, True)
        logits64 = port.head_forward({k: v.clone() for k, v in st64.items()}, feat64, 1.0, True)
    model = model.to(DEV).train()
    with dp_b200.compute_mode("bf16"), torch.no_grad():
        feat = model.res2plus1d(xd)
        logits = model.linear(feat)
    with torch.no_grad():
        logits_h = port.head_forward({k: v.clone() for k, v in st64.items()}, feat.double(), 1.0, True)
    e_feat = ((feat.double() - feat64).norm() / feat64.norm()).item()
    lmax = logits64.abs().max()
    e_logit = ((logits.double() - logits64).abs().max() / lmax).item()
    e_head = ((logits.double() - logits_h).abs().max() / lmax).item()
    print(f"features rel-L2 {e_feat:.2e}; logits vs fp64 {e_logit:.2e}; logits vs fp64 head on the CUDA features {e_head:.2e}")
    assert e_feat < 2e-2
    assert e_head < 1e-3            # the head itself (fp32 PyTorch ops) is exact: all of e_logit is amplified feature error
    assert e_logit < 0.3


@pytest.mark.parametrize("tag,alpha,loss_name", [("a1.0_focal", 1.0, "focal"), ("a1.0_ldam", 1.0, "ldam"), ("a0.01_focal", 0.01, "focal")])
def test_baseline_config_golden_step(golden_dir, tag, alpha, loss_name):
    """The BASELINE.json configs[1] model at full resolution, one training step on 4 seeded clips, against the golden
    vectors of the unmodified reference (oracle/make_golden_r2.py): fp32 validation mode at north_star's 1e-4."""
    gold = np.load(os.path.join(golden_dir, "baseline_config_step.npz"))
    x, y = port.synthetic_clips(4, 21, 128, 128)
    y = torch.from_numpy(gold["y"])
    w = dp_b200.rw_class_weights([300, 17000]).to(DEV)
    ref_logits, ref_loss = torch.from_numpy(gold[tag + "_logits"]), float(gold[tag + "_loss"])
    for mode in ("fp32", "bf16"):
        torch.manual_seed(42)
        m = R2Plus1DClassifier((3, 21, 128, 128), 2, [1, 2, 2, 1], False, alpha).to(DEV).train()
        lf = FocalLoss(weight=w, gamma=2.0) if loss_name == "focal" else LDAMLoss([300, 17000], max_m=0.5, weight=w, s=1.0)
        with dp_b200.compute_mode(mode):
            logits = m(x.to(DEV))
            loss = lf(logits, y.to(DEV))
            loss.backward()
        e_logit = ((logits.detach().cpu() - ref_logits).abs().max() / ref_logits.abs().max()).item()
        e_loss = abs(loss.item() - ref_loss) / abs(ref_loss)
        print(f"[{tag} {mode}] logits rel {e_logit:.2e} loss rel {e_loss:.2e}")
        if mode == "fp32":
            assert e_logit < 1e-4 and e_loss < 1e-4
            params = dict(m.named_parameters())
            gn = gold[tag + "_grad_norm"]
            worst = 0.0
            for i, n in enumerate(gold[tag + "_grad_names"]):
                g = params[str(n)].grad
                if gn[i] < 1e-6 * gn.max():
                    continue
                worst = max(worst, abs(g.double().norm().item() - gn[i]) / gn[i])
            print(f"[{tag} fp32] worst per-parameter gradient-norm deviation {worst:.2e}")
            assert worst < 5e-2     # ill-conditioned sums: the fp32 reference is itself ~1e-2 from fp64
            sd = m.state_dict()
            from tests.test_gpu_slowfast import summarise
            for k, want in zip(gold[tag + "_bn_keys"], gold[tag + "_bn_summary"]):
                got = summarise(sd[str(k)])
                assert np.allclose(got[:2], want[:2], rtol=2e-3, atol=2e-3), (k, got, want)
        else:
            assert torch.isfinite(loss)
            assert e_logit < 0.3 and e_loss < 0.15      # bf16 storage envelope at batch 4


def test_slowfast_every_conv_takes_tcgen05():
    """All 74 convolutions of SlowFast (channels 4..512) -- forward, data gradient, weight gradient -- run on the
    tcgen05 family in bf16: with the `strict_tc` option a planner refusal raises instead of falling back, and the
    CUDA-core launch counters stay at zero."""
    from dp_b200.slowfast import Bottleneck3D, SlowFast
    lib = _lib.load()
    torch.manual_seed(42)
    m = SlowFast((3, 20, 128, 128), Bottleneck3D, [1, 2, 2, 1], 4, 1, 2, 1.0).to(DEV).train()
    g = torch.Generator().manual_seed(1)
    x = (torch.randint(0, 256, (2, 3, 20, 128, 128), generator=g).float() - 96.0).to(DEV)
    y = torch.tensor([0, 1], device=DEV)
    s0, f0 = lib.dp_simt_launch_count(), lib.dp_simt_fallback_count()
    _lib.set_option("strict_tc", 1)
    try:
        with dp_b200.compute_mode("bf16", "auto"):
            loss = FocalLoss(weight=torch.ones(2, device=DEV))(m(x), y)
            loss.backward()
        torch.cuda.synchronize()
    finally:
        _lib.set_option("strict_tc", 0)
    assert torch.isfinite(loss)
    assert lib.dp_simt_launch_count() - s0 == 0 and lib.dp_simt_fallback_count() - f0 == 0
    n_conv = sum(1 for mod in m.modules() if isinstance(mod, torch.nn.Conv3d))
    print(f"{n_conv} Conv3d modules, 0 CUDA-core launches")
