"""GPU: the callers either side of the hot path -- the sliding-window inference loop (config 5), the
uint8 input boundary (n3) and the whole-step CUDA graph -- against the oracle / the eager path."""
import os
import sys

import pytest
import torch

import dp_b200
from dp_b200 import inference
from dp_b200.R2Plus1D import R2Plus1DClassifier
from dp_b200.loss import FocalLoss
from dp_b200.optim import FusedClipAdamW

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import r2plus1d_port as port  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _trained_state(layer_sizes, alpha, T, H, W):
    """Seeded model whose BatchNorm running statistics have seen one batch (eval mode needs non-trivial stats)."""
    torch.manual_seed(42)
    model = R2Plus1DClassifier((3, T, H, W), 2, layer_sizes, False, alpha)
    st = port.clone_state({k: v.clone() for k, v in model.state_dict().items()}, requires_grad=False)
    xs, _ = port.structured_clips(4, T, H, W, seed=77)
    with torch.no_grad():
        port.classifier_forward(st, xs, layer_sizes, alpha, training=True)
    model.load_state_dict({k: v.detach() for k, v in st.items()})
    return model, st


def test_sliding_window_matches_reference_loop():
    """`generate_prob_curve` loop (/root/reference/src/utils/utility.py:936-949): window i = frames i+1..i+seq_len,
    eval mode, softmax[:,0]; here batched, uint8 frames resident on the device, optionally sharded by window range."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    layer_sizes, alpha, seq_len, dist, H, W = [1, 1, 1, 1], 0.01, 21, 3, 64, 64
    model, st = _trained_state(layer_sizes, alpha, seq_len, H, W)
    g = torch.Generator().manual_seed(4321)
    n_frames = 40
    base, _ = port.structured_clips(1, n_frames, H, W, seed=5)                       # one "shot": (1,3,N,H,W) minus mean
    mean = torch.tensor([90.0, 98.0, 102.0]).view(1, 3, 1, 1, 1)
    frames = (base + mean).round().clamp(0, 255).to(torch.uint8)[0].permute(1, 2, 3, 0).contiguous()  # (N,H,W,3)
    n_win = inference.num_windows(n_frames, seq_len, dist)
    assert n_win == 16
    ref = []
    for i in range(n_win):                                                           # the reference's batch-1 loop
        clip = frames[i + 1:i + 1 + seq_len].float() - torch.tensor([90.0, 98.0, 102.0])
        x = clip.permute(3, 0, 1, 2).unsqueeze(0)                                    # (1,3,T,H,W)
        with torch.no_grad():
            out = port.classifier_forward(st, x, layer_sizes, alpha, training=False)
        ref.append(torch.softmax(out, 1)[0, 0].item())
    ref = torch.tensor(ref)
    model = model.to(DEV).eval()
    fr = frames.to(DEV)
    for mode, tol in (("fp32", 1e-4), ("bf16", 5e-2)):
        with dp_b200.compute_mode(mode):
            p8 = inference.sliding_window_probs(model, fr, seq_len, dist, batch_size=8).cpu()
            p1 = inference.sliding_window_probs(model, fr, seq_len, dist, batch_size=1).cpu()
            shards = [inference.sliding_window_probs(model, fr, seq_len, dist, batch_size=5,
                                                     window_range=dp_b200.distributed.shard_range(n_win, r, 3)).cpu()
                      for r in range(3)]
            # the same windows without the CUDA-graph replay, without the per-frame stem cache, and after a weight change
            pe = inference.sliding_window_probs(model, fr, seq_len, dist, batch_size=8, use_graph=False).cpu()
            pn = inference.sliding_window_probs(model, fr, seq_len, dist, batch_size=8, stem_cache=False).cpu()
        assert torch.allclose(pe, p8, atol=1e-6), mode                  # graph replay == eager launches, bit for bit up to softmax
        assert (pn - ref).abs().max().item() < tol
        assert p8.shape == (n_win,)
        print(f"[{mode}] max |dP| vs reference loop: {(p8 - ref).abs().max().item():.3e}")
        assert (p8 - ref).abs().max().item() < tol
        assert (p1 - ref).abs().max().item() < tol
        assert torch.allclose(torch.cat(shards), p8, atol=1e-6 if mode == "fp32" else 2e-2)
    assert model.training is False
    # captured graphs hold packed weight copies: a weight update must invalidate them
    with dp_b200.compute_mode("bf16"):
        before = inference.sliding_window_probs(model, fr, seq_len, dist, batch_size=8).cpu()
        with torch.no_grad():
            model.linear[3].bias.add_(torch.tensor([0.5, -0.5], device=DEV))
            model.res2plus1d.conv5.block1.conv2.temporal_conv.conv.weight.mul_(1.5)
        after_graph = inference.sliding_window_probs(model, fr, seq_len, dist, batch_size=8).cpu()
        after_eager = inference.sliding_window_probs(model, fr, seq_len, dist, batch_size=8, use_graph=False).cpu()
    assert (before - after_graph).abs().max().item() > 1e-3
    assert torch.allclose(after_graph, after_eager, atol=1e-6)
    curve = inference.postprocess_curve(ref.tolist(), clip_len=seq_len, frame_srt=2)
    assert len(curve) == seq_len + 2 + n_win - 2


def test_uint8_frames_equal_float_clips():
    """The uint8 input boundary (mean-subtract + layout fused into the stem) gives what the float NCDHW clip gives."""
    layer_sizes, alpha, T, H, W = [1, 1, 1, 1], 0.01, 9, 64, 64
    model, _ = _trained_state(layer_sizes, alpha, T, H, W)
    model = model.to(DEV).eval()
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (3, T, H, W, 3), generator=g, dtype=torch.uint8).to(DEV)
    x = (frames.float() - torch.tensor([90.0, 98.0, 102.0], device=DEV)).permute(0, 4, 1, 2, 3).contiguous()
    for mode in ("bf16", "fp32"):
        with dp_b200.compute_mode(mode), torch.no_grad():
            a = model.res2plus1d(frames, (90.0, 98.0, 102.0))
            b = model.res2plus1d(x)
        assert torch.equal(a, b), mode


def test_graphed_step_equals_eager_steps():
    """Whole-step CUDA graph replay (dp_b200.graph.GraphedTrainStep) reproduces the eager step sequence."""
    layer_sizes, alpha, B, T, H, W = [1, 1, 1, 1], 0.01, 4, 9, 64, 64
    x, y = port.structured_clips(B, T, H, W, seed=9)
    y[0], y[1] = 0, 1
    x, y = x.to(DEV), y.to(DEV)
    w = dp_b200.rw_class_weights([300, 17000]).to(DEV)

    def make():
        torch.manual_seed(42)
        m = R2Plus1DClassifier((3, T, H, W), 2, layer_sizes, False, alpha).to(DEV).train()
        return m, FocalLoss(weight=w, gamma=2.0), FusedClipAdamW(m.parameters(), lr=1e-3, max_norm=1.0, capturable=True)

    n_steps = 5
    m1, lf1, o1 = make()
    eager = []
    for _ in range(n_steps):
        o1.zero_grad(set_to_none=True)
        loss = lf1(m1(x), y)
        loss.backward()
        o1.step()
        eager.append(loss.item())
    del loss
    m2, lf2, o2 = make()
    from dp_b200.graph import GraphedTrainStep
    g = GraphedTrainStep(m2, lf2, o2, x, y, warmup=1)      # one eager warm-up step; capture records, it does not run
    graphed = [g.step(x, y)[0].item() for _ in range(n_steps - 1)]
    print("eager ", eager)
    print("graph ", graphed)
    assert g.launches_per_step > 50
    for a, b in zip(eager[1:], graphed):
        assert abs(a - b) <= 2e-3 * max(abs(a), 1e-6), (eager, graphed)
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(p1, p2, rtol=1e-3, atol=1e-5), n
