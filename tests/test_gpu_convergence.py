"""GPU: >= 300-step convergence A/B of the bf16 product mode against the fp32 validation mode and the oracle port on a
learnable synthetic task, and >= 99.9 % thresholded-label agreement on 1024 held-out clips (tests/convergence_ab.py;
north_star's label criterion, /root/reference/src/evaluate.py:56-57).  This is the evidence beside the re-scoped bf16
logit tolerance."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu


def test_bf16_converges_like_fp32_and_the_port():
    from tests.convergence_ab import ROOT, check, run_ab
    res = run_ab(steps=300, batch=16, heldout=1024)
    ok = check(res)
    res["within_band"] = bool(ok)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(res, open(os.path.join(out, "convergence_ab.json"), "w"))
    print(json.dumps({k: v for k, v in res.items() if k not in ("loss", "acc")}, indent=1))
    assert res["heldout_same_weights"]["label_agreement_bf16_vs_port"] >= 0.999
    assert res["heldout_same_weights"]["label_agreement_fp32_vs_port"] >= 0.999
    assert ok, "loss curves / accuracies left the stated band"
