"""CPU: the oracle port (oracle/r2plus1d_port.py) against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py imported /root/reference in the build container; fixtures in tests/golden/).

The reference's own tests hold no golden vectors for this path, so this file is what
pins the oracle: same seeded state + same synthetic clips must give the reference's logits, losses,
gradients, BatchNorm running statistics and eval-mode logits; the loss formulas must reproduce the
reference's values and logit gradients on random logits; RW / DRW weights must match to the last bit.
"""
import os
import sys

import numpy as np
import pytest
import torch

import dp_b200
from dp_b200.R2Plus1D import R2Plus1DClassifier

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import r2plus1d_port as port  # noqa: E402
from oracle import np_ops  # noqa: E402

CLS = [300, 17000]


def summarise(t):
    f = t.detach().double().reshape(-1)
    head = torch.zeros(4, dtype=torch.float64)
    head[:min(4, f.numel())] = f[:4]
    return np.concatenate([[f.sum().item(), f.abs().sum().item()], head.numpy()])


def seeded_state(input_size, layer_sizes, alpha, seed=42):
    """The drop-in classifier constructs its sub-modules in the reference's order, so torch seed 42 yields
    the reference's initial weights (checked against init_seed42.npz below)."""
    torch.manual_seed(seed)
    m = R2Plus1DClassifier(input_size, 2, layer_sizes, False, alpha)
    return {k: v.clone() for k, v in m.state_dict().items()}


def test_seed42_init_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "init_seed42.npz"))
    sd = seeded_state((3, 21, 128, 128), [1, 2, 2, 1], 1.0)
    assert list(sd.keys()) == [str(k) for k in gold["keys"]]
    assert len(sd) == 201
    for k, shape, ref in zip(gold["keys"], gold["shapes"], gold["summary"]):
        t = sd[str(k)]
        assert str(tuple(t.shape)) == str(shape), k
        np.testing.assert_allclose(summarise(t), ref, rtol=1e-6, atol=1e-9, err_msg=str(k))
    n_params = sum(v.numel() for k, v in sd.items() if "running_" not in k and "num_batches" not in k)
    assert n_params == 1_587_523


@pytest.mark.parametrize("alpha", [1.0, 0.01])
@pytest.mark.parametrize("loss_name", ["focal", "ldam", "ce"])
def test_port_train_step_matches_reference(golden_dir, alpha, loss_name):
    gold = np.load(os.path.join(golden_dir, "small_train_step.npz"))
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    B, T, H, W = 4, 5, 32, 32
    x, y = port.synthetic_clips(B, T, H, W)
    y[0], y[1] = 0, 1
    np.testing.assert_allclose(summarise(x), gold["x_summary"], rtol=0, atol=0)
    assert np.array_equal(y.numpy(), gold["y"])
    layer_sizes = [1, 1, 1, 1]
    st = port.clone_state(seeded_state((3, T, H, W), layer_sizes, alpha))
    w = port.rw_class_weights(CLS)
    logits, loss, grads = port.train_step(st, x, y, layer_sizes, alpha, loss=loss_name, weight=w,
                                          margins=port.ldam_margins(CLS, 0.5), s=1.0)
    tag = f"a{alpha}_{loss_name}"
    np.testing.assert_allclose(logits.numpy(), gold[tag + "_logits"], rtol=2e-5, atol=2e-6)
    assert abs(loss.item() - float(gold[tag + "_loss"])) <= 2e-5 * abs(float(gold[tag + "_loss"]))
    gn = gold[tag + "_grad_norm"]
    for i, n in enumerate(gold[tag + "_grad_names"]):
        g = grads[str(n)]
        assert abs(g.double().norm().item() - gn[i]) <= 1e-3 * max(gn[i], 1e-3 * gn.max()), n
    for k, ref in zip(gold[tag + "_bn_keys"], gold[tag + "_bn_summary"]):
        np.testing.assert_allclose(summarise(st[str(k)]), ref, rtol=1e-4, atol=1e-6, err_msg=str(k))
    with torch.no_grad():
        ev = port.classifier_forward(st, x, layer_sizes, alpha, training=False)
    np.testing.assert_allclose(ev.numpy(), gold[tag + "_eval_logits"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("C_", [2, 5])
def test_port_losses_match_reference(golden_dir, C_):
    gold = np.load(os.path.join(golden_dir, "loss_kat.npz"))
    target = torch.from_numpy(gold[f"c{C_}_target"])
    w = torch.from_numpy(gold[f"c{C_}_weight"])
    counts = gold[f"c{C_}_counts"].tolist()
    cases = {
        "focal_g2": lambda z: port.focal_loss(z, target, w, 2.0),
        "focal_g0": lambda z: port.focal_loss(z, target, w, 0.0),
        "focal_g1.5": lambda z: port.focal_loss(z, target, w, 1.5),
        "ldam_s30": lambda z: port.ldam_loss(z, target, port.ldam_margins(counts, 0.5), w, 30.0),
        "ldam_s1": lambda z: port.ldam_loss(z, target, port.ldam_margins(counts, 0.5), w, 1.0),
        "ldam_s1_now": lambda z: port.ldam_loss(z, target, port.ldam_margins(counts, 0.5), None, 1.0),
        "ce": lambda z: port.ce_loss(z, target, w),
        "ce_now": lambda z: port.ce_loss(z, target, None),
    }
    for name, fn in cases.items():
        z = torch.from_numpy(gold[f"c{C_}_logits"]).clone().requires_grad_(True)
        val = fn(z)
        val.backward()
        ref = float(gold[f"c{C_}_{name}_loss"])
        assert abs(val.item() - ref) <= 2e-6 * max(1.0, abs(ref)), name
        gref = gold[f"c{C_}_{name}_grad"]
        atol = 1e-5 * max(1.0, float(np.abs(gref).max()))    # s=30 scales the fp32 rounding of the reference itself
        np.testing.assert_allclose(z.grad.numpy(), gref, rtol=2e-5, atol=atol, err_msg=name)
        if name.startswith("ldam"):
            np.testing.assert_array_equal(port.ldam_margins(counts, 0.5).numpy(), gold[f"c{C_}_{name}_m"])
        # the independent numpy restatement agrees too (float64)
        kind = name.split("_")[0]
        kw = dict(gamma=float(name.split("_g")[1]) if kind == "focal" else 2.0,
                  s=30.0 if "s30" in name else 1.0,
                  margins=port.ldam_margins(counts, 0.5).numpy().astype(np.float64) if kind == "ldam" else None,
                  weight=None if name.endswith("_now") else w.numpy().astype(np.float64))
        lv, lg = np_ops.loss_and_grad(kind, gold[f"c{C_}_logits"].astype(np.float64), target.numpy(), **kw)
        assert abs(lv - ref) <= 2e-6 * max(1.0, abs(ref)), name
        np.testing.assert_allclose(lg, gref, rtol=2e-5, atol=atol, err_msg=name)


def test_loss_closed_forms():
    """SURVEY.md section 8(a) known-answer formulas: CE = lse(z) - z_y; Focal dL/dCE; LDAM margins."""
    z = np.array([[0.3, -1.2], [2.0, 0.5], [-0.7, -0.1]])
    y = np.array([0, 1, 1])
    w = np.array([0.98, 0.02])
    ce = np.log(np.exp(z).sum(1)) - z[np.arange(3), y]
    p = np.exp(-ce)
    lv, lg = np_ops.loss_and_grad("focal", z, y, weight=w, gamma=2.0)
    assert abs(lv - (w[y] * (1 - p) ** 2 * ce).sum()) < 1e-12
    dce = w[y] * ((1 - p) ** 2 + 2 * (1 - p) * p * ce)
    sm = np.exp(z) / np.exp(z).sum(1, keepdims=True)
    onehot = np.eye(2)[y]
    np.testing.assert_allclose(lg, dce[:, None] * (sm - onehot), atol=1e-12)
    m = port.ldam_margins([300, 17000], 0.5).numpy()
    np.testing.assert_allclose(m, 0.5 * np.array([1.0, (300 / 17000) ** 0.25]), rtol=1e-6)


def test_class_weights_match_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "class_weights.npz"))
    for fn in (port.rw_class_weights, dp_b200.rw_class_weights):
        np.testing.assert_array_equal(fn(CLS).numpy(), gold["rw"])
        np.testing.assert_array_equal(fn(CLS, False).numpy(), np.array([1.0, 1.0], dtype=np.float32))
    betas = dp_b200.drw_betas(0.25)
    assert betas == [0, 0.25, 0.5, 0.75]
    for fn in (port.drw_class_weights, dp_b200.drw_class_weights):
        for epoch in (0, 31, 32, 63, 64, 95, 96, 127):
            np.testing.assert_array_equal(fn(epoch, 128, betas, CLS).numpy(), gold[f"drw_e{epoch}"])
        for epoch in (0, 32, 64, 96):
            np.testing.assert_array_equal(fn(epoch, 128, betas, [3, 40]).numpy(), gold[f"drw_small_e{epoch}"])


@pytest.mark.parametrize("geom", [
    # C, K, kernel, stride, padding, (T, H, W)
    (3, 5, (1, 7, 7), (1, 2, 2), (0, 3, 3), (3, 12, 12)),
    (6, 4, (3, 1, 1), (2, 1, 1), (1, 0, 0), (5, 4, 4)),
    (4, 7, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 6, 5)),
    (5, 3, (1, 1, 1), (1, 2, 2), (0, 0, 0), (3, 7, 7)),
])
def test_numpy_primitives_agree_with_port(geom):
    """oracle/np_ops.py (pure numpy loops, float64) against the torch primitives the port calls:
    conv3d forward / dgrad / wgrad, train-mode BatchNorm forward / backward, LeakyReLU."""
    C_, K, k, s, p, (T, H, W) = geom
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, C_, T, H, W))
    w = rng.standard_normal((K, C_, *k))
    xt = torch.from_numpy(x).requires_grad_(True)
    wt = torch.from_numpy(w).requires_grad_(True)
    yt = torch.nn.functional.conv3d(xt, wt, None, s, p)
    y = np_ops.conv3d(x, w, s, p)
    np.testing.assert_allclose(y, yt.detach().numpy(), atol=1e-10)
    dy = rng.standard_normal(y.shape)
    yt.backward(torch.from_numpy(dy))
    np.testing.assert_allclose(np_ops.conv3d_dgrad(dy, w, x.shape, s, p), xt.grad.numpy(), atol=1e-10)
    np.testing.assert_allclose(np_ops.conv3d_wgrad(x, dy, w.shape, s, p), wt.grad.numpy(), atol=1e-9)
    # BatchNorm (train) + LeakyReLU
    gamma, beta = rng.standard_normal(K), rng.standard_normal(K)
    yt2 = torch.from_numpy(y).requires_grad_(True)
    gt, bt = torch.from_numpy(gamma).requires_grad_(True), torch.from_numpy(beta).requires_grad_(True)
    rm, rv = torch.zeros(K, dtype=torch.float64), torch.ones(K, dtype=torch.float64)
    zt = torch.nn.functional.leaky_relu(torch.nn.functional.batch_norm(yt2, rm, rv, gt, bt, True, 0.1, 1e-5), 0.01)
    z, cache = np_ops.bn_lrelu_fwd(y, gamma, beta, 0.01, 1e-5)
    np.testing.assert_allclose(z, zt.detach().numpy(), atol=1e-10)
    np.testing.assert_allclose(cache["running_mean"], rm.numpy(), atol=1e-12)
    np.testing.assert_allclose(cache["running_var"], rv.numpy(), atol=1e-12)
    dz = rng.standard_normal(z.shape)
    zt.backward(torch.from_numpy(dz))
    dyn, dg, db = np_ops.bn_lrelu_bwd(dz, cache)
    np.testing.assert_allclose(dyn, yt2.grad.numpy(), atol=1e-9)
    np.testing.assert_allclose(dg, gt.grad.numpy(), atol=1e-9)
    np.testing.assert_allclose(db, bt.grad.numpy(), atol=1e-9)


def test_bf16_storage_emulation_is_close_to_fp32_port():
    """The bf16-storage variant of the port is the same algorithm with roundings inserted; on a small case
    it must stay within a few bf16 ulps of the fp32 port and produce finite gradients."""
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    layer_sizes, alpha = [1, 1, 1, 1], 0.01
    state = seeded_state((3, 5, 32, 32), layer_sizes, alpha)
    x, y = port.structured_clips(4, 5, 32, 32)
    y[0], y[1] = 0, 1
    w = port.rw_class_weights(CLS)
    l32, loss32, _ = port.train_step(port.clone_state(state), x, y, layer_sizes, alpha, "focal", w)
    l16, loss16, g16 = port.train_step(port.clone_state(state), x, y, layer_sizes, alpha, "focal", w, storage="bf16")
    assert ((l16 - l32).abs().max() / l32.abs().max()).item() < 0.1
    assert abs(loss16.item() - loss32.item()) / abs(loss32.item()) < 0.1
    assert all(torch.isfinite(g).all() for g in g16.values())
