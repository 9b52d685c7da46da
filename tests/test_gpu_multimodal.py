"""GPU: the R(2+1)D-backed multimodal model + GradientBlending (config 4), part-wise (the
reference has no R(2+1)D multimodal model to compare the whole against)."""
import os
import sys

import pytest
import torch

import dp_b200
from dp_b200.MultiModal import GradientBlending, MultiModalR2Plus1D, MultiModalR2Plus1D_GB
from dp_b200.loss import CELoss, FocalLoss, LDAMLoss

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import r2plus1d_port as port  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"
ARGS_V = {"layer_sizes": [1, 1, 1, 1], "alpha": 0.01}
ARGS_T = dict(n_features=18, kernel_size=5, feature_dims=128, max_len=21, n_layers=2, n_heads=8, dim_feedforward=256,
              dropout=0.0)


def _data(B=4, T=9, H=64, W=64):
    x, y = port.structured_clips(B, T, H, W, seed=21)
    y[0], y[1] = 0, 1
    g = torch.Generator().manual_seed(8)
    ts = torch.randn(B, 21, 18, generator=g)
    return x.to(DEV), ts.to(DEV), y.to(DEV)


def test_fusion_model_video_branch_matches_oracle_encoder():
    torch.manual_seed(42)
    m = MultiModalR2Plus1D(2, ARGS_V, ARGS_T).to(DEV).train()
    x, ts, y = _data()
    st = {"res2plus1d." + k: v.detach().cpu().clone() for k, v in m.encoder_video.state_dict().items()}
    with torch.no_grad():
        ref = port.encoder_forward(st, x.cpu(), ARGS_V["layer_sizes"], ARGS_V["alpha"], training=True)
    with dp_b200.compute_mode("fp32"), torch.no_grad():
        feat = m.encoder_video(x)
    assert ((feat.cpu() - ref).abs().max() / ref.abs().max()).item() < 1e-4
    with dp_b200.compute_mode("bf16"):
        out = m(x, ts)
        loss = FocalLoss(weight=torch.ones(2, device=DEV))(out, y)
        loss.backward()
    assert out.shape == (4, 2) and torch.isfinite(loss)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    h, hv, ht = m.encode(x, ts)
    assert h.shape == (4, (128 + 128) // 2) and hv.shape == (4, 128) and ht.shape == (4, 128)


def test_gradient_blending_three_heads():
    torch.manual_seed(7)
    m = MultiModalR2Plus1D_GB(2, ARGS_V, ARGS_T, use_stream="multi-GB").to(DEV).train()
    m.ts_model.encoder.noise.eval()          # deterministic 0D branch
    x, ts, y = _data()
    w = dp_b200.rw_class_weights([300, 17000]).to(DEV)
    gb = GradientBlending(FocalLoss(weight=w, gamma=2.0), LDAMLoss([300, 17000], 0.5, w, s=1.0), CELoss(weight=w),
                          vis_weight=0.1, ts_weight=0.4, vis_ts_weight=0.5)
    with dp_b200.compute_mode("fp32"):
        out_multi, out_vis, out_ts = m(x, ts)
        loss = gb(out_multi, out_vis, out_ts, y)
        loss.backward()
    # the blended loss against the oracle's loss formulas on the same logits
    wc, yc = w.cpu(), y.cpu()
    want = (0.1 * port.focal_loss(out_vis.detach().cpu(), yc, wc, 2.0) +
            0.4 * port.ldam_loss(out_ts.detach().cpu(), yc, port.ldam_margins([300, 17000], 0.5), wc, 1.0) +
            0.5 * port.ce_loss(out_multi.detach().cpu(), yc, wc))
    assert abs(loss.item() - want.item()) <= 1e-5 * abs(want.item())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # stream switches (reference MultiModal.py:136-154)
    m.update_use_stream("video")
    with dp_b200.compute_mode("bf16"):
        assert m(x, ts).shape == (4, 2)
    m.update_use_stream("0D")
    assert m(x, ts).shape == (4, 2)
    m.update_use_stream("multi")
    with dp_b200.compute_mode("bf16"):
        assert m(x, ts).shape == (4, 2)
    gb.update_weights({"video": 0.2, "0D": 0.3, "multi": 0.5})
    assert (gb.vis_weight, gb.ts_weight, gb.vis_ts_weight) == (0.2, 0.3, 0.5)


def test_video_only_weighting_reduces_to_the_classifier_step():
    """With w_ts = w_multi = 0 the video branch receives exactly the gradient of the stand-alone classifier."""
    from dp_b200.R2Plus1D import R2Plus1DClassifier
    torch.manual_seed(11)
    m = MultiModalR2Plus1D_GB(2, ARGS_V, ARGS_T).to(DEV).train()
    m.ts_model.encoder.noise.eval()
    x, ts, y = _data()
    w = torch.ones(2, device=DEV)
    c = R2Plus1DClassifier((3, 9, 64, 64), 2, ARGS_V["layer_sizes"], False, ARGS_V["alpha"]).to(DEV).train()
    c.res2plus1d.load_state_dict(m.vis_model.res2plus1d.state_dict())
    c.linear.load_state_dict(m.vis_model.linear.state_dict())
    gb = GradientBlending(FocalLoss(weight=w), FocalLoss(weight=w), FocalLoss(weight=w), 1.0, 0.0, 0.0)
    with dp_b200.compute_mode("fp32"):
        o = m(x, ts)
        gb(*o, y).backward()
        FocalLoss(weight=w)(c(x), y).backward()
    for (n, p), (_, q) in zip(m.vis_model.res2plus1d.named_parameters(), c.res2plus1d.named_parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-4, atol=1e-7), n
