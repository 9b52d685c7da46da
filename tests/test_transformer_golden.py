"""CPU: the 0D branch of the multimodal model (MultiModal.TransformerEncoder) against golden vectors produced by the
UNMODIFIED reference TransformerEncoder (/root/reference/src/models/transformer.py:39-138; oracle/make_golden_r2.py).
The branch is plain PyTorch in both (it is not on the CUDA hot path), so the check runs on the CPU."""
import os

import numpy as np
import torch

from dp_b200.MultiModal import TransformerEncoder

T_ARGS = dict(n_features=18, kernel_size=5, feature_dims=128, max_len=21, n_layers=2, n_heads=8, dim_feedforward=256, dropout=0.0)


def summarise(t):
    f = t.detach().double().reshape(-1)
    head = torch.zeros(4, dtype=torch.float64)
    head[:min(4, f.numel())] = f[:4]
    return np.concatenate([[f.sum().item(), f.abs().sum().item()], head.numpy()])


def test_transformer_encoder_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "transformer_encoder.npz"))
    torch.manual_seed(42)
    m = TransformerEncoder(**T_ARGS)
    sd = m.state_dict()
    # same keys, shapes and -- same construction order under the same seed -- same initial values
    assert list(sd.keys()) == [str(k) for k in gold["keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in gold["shapes"]]
    got = np.stack([summarise(v.float()) for v in sd.values()])
    np.testing.assert_allclose(got, gold["summary"], rtol=1e-6, atol=1e-7)
    x = torch.from_numpy(gold["x"])
    m.eval()
    with torch.no_grad():
        ev = m(x)
    np.testing.assert_allclose(ev.numpy(), gold["eval_out"], rtol=1e-5, atol=1e-6)
    m.train()
    m.noise.eval()
    o = m(x)
    (o * o).sum().backward()
    np.testing.assert_allclose(o.detach().numpy(), gold["train_out"], rtol=1e-5, atol=1e-6)
    params = dict(m.named_parameters())
    gmax = float(gold["grad_norm"].max())
    for n, want in zip(gold["grad_names"], gold["grad_norm"]):
        if want < 1e-5 * gmax:      # a bias in front of a BatchNorm: its gradient is rounding noise in the reference too
            continue
        g = params[str(n)].grad
        assert g is not None, n
        assert abs(g.double().norm().item() - want) <= 1e-4 * max(want, 1e-6) + 1e-7, (n, g.double().norm().item(), want)
