"""CPU: the C-ABI library loads and exports every symbol include/dp_b200.h declares (no compute calls), the
ctypes prototypes agree with the header, the drop-in modules keep the reference's surface, and the
product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import dp_b200
from dp_b200 import _lib, distributed as dpd, inference
from dp_b200.R2Plus1D import Conv3dBlock, R2Plus1DClassifier, SpatioTemporalConv, SpatioTemporalResBlock
from dp_b200.loss import CELoss, FocalLoss, ImbalancedDatasetSampler, LDAMLoss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dp_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    # "<ret> dp_xxx(" at the start of a declaration
    return sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libdp_b200.so must be built in-tree (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dp_b200.h but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_header_argument_counts_match_binding():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), (name, n, len(args))


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.dp_version() >= 100
    assert lib.dp_launch_count() == 0
    assert _lib.get_option("tc_enable") in (0, 1)
    with pytest.raises(_lib.DpError):
        _lib.set_option("no_such_option", 1)
    # struct layout mirrors dp_conv_desc: 21 int32
    assert ctypes.sizeof(_lib.ConvDesc) == 21 * 4


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = R2Plus1DClassifier((3, 5, 32, 32), 2, [1, 1, 1, 1], False, 0.01)
    with pytest.raises(_lib.DpError):
        m(torch.zeros(2, 3, 5, 32, 32))
    with pytest.raises(_lib.DpError):
        FocalLoss(weight=torch.ones(2))(torch.zeros(4, 2), torch.zeros(4, dtype=torch.long))
    with pytest.raises(_lib.DpError):
        _lib.require_device()


def test_module_surface_matches_reference():
    """State-dict keys, shapes, attribute names and constructor defaults the reference's callers rely on
    (SURVEY.md section 8b)."""
    m = R2Plus1DClassifier((3, 21, 128, 128), 2, [1, 2, 2, 1], False, 1.0)
    sd = m.state_dict()
    assert len(sd) == 201
    assert sd["res2plus1d.conv1.spatio_conv.conv.weight"].shape == (45, 3, 1, 7, 7)
    assert sd["res2plus1d.conv1.temporal_conv.conv.weight"].shape == (32, 45, 3, 1, 1)
    assert sd["res2plus1d.conv3.block1.downsample_conv.spatio_conv.conv.weight"].shape == (21, 32, 1, 1, 1)
    assert sd["res2plus1d.conv5.block1.conv2.spatio_conv.conv.weight"].shape == (288, 128, 1, 3, 3)
    assert "res2plus1d.conv2.block1.conv1.spatio_conv.bn.num_batches_tracked" in sd
    assert sd["linear.3.weight"].shape == (2, 64)
    assert m.input_size == (3, 21, 128, 128)
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "pool"):
        assert hasattr(m.res2plus1d, name)
    mids = [m.res2plus1d.conv2.block1.conv1.spatio_conv.conv.out_channels,
            m.res2plus1d.conv3.block1.conv1.spatio_conv.conv.out_channels,
            m.res2plus1d.conv3.block1.conv2.spatio_conv.conv.out_channels,
            m.res2plus1d.conv5.block1.conv1.spatio_conv.conv.out_channels,
            m.res2plus1d.conv5.block1.conv2.spatio_conv.conv.out_channels,
            m.res2plus1d.conv3.block1.downsample_conv.spatio_conv.conv.out_channels,
            m.res2plus1d.conv5.block1.downsample_conv.spatio_conv.conv.out_channels]
    assert mids == [72, 115, 144, 230, 288, 21, 42]
    # inner activations use the SpatioTemporalConv default slope, block end / stem use alpha (R2Plus1D.py:172-179,210)
    assert m.res2plus1d.conv2.block1.conv1.spatio_conv.relu.negative_slope == 0.01
    assert m.res2plus1d.conv2.block1.relu.negative_slope == 1.0
    assert m.res2plus1d.conv1.spatio_conv.relu.negative_slope == 1.0
    # int kernel / stride / padding promotion (R2Plus1D.py:29-42)
    b = Conv3dBlock(4, 8, 3, 2, 1, 1)
    assert b.conv.kernel_size == (1, 3, 3) and b.conv.stride == (1, 2, 2) and b.conv.padding == (0, 1, 1)
    s = SpatioTemporalConv(32, 64, (3, 3, 3), (2, 2, 2), 1, (1, 1, 1))
    assert s.spatio_conv.conv.out_channels == 115 and s.temporal_conv.conv.stride == (2, 1, 1)
    r = SpatioTemporalResBlock(32, 64, 3, downsample=True)
    assert r.downsample_conv.spatio_conv.conv.kernel_size == (1, 1, 1)
    with pytest.raises(NotImplementedError):
        Conv3dBlock(4, 8, 3, 1, 2, 1)
    # the optimiser / clip / checkpoint protocol sees ordinary fp32 parameters
    assert all(p.dtype == torch.float32 for p in m.parameters())
    m2 = R2Plus1DClassifier((3, 21, 128, 128), 2, [1, 2, 2, 1], False, 1.0)
    m2.load_state_dict(sd)
    with pytest.raises(ValueError):
        R2Plus1DClassifier((3, 21, 4, 4), 2, [1, 1, 1, 1])._check_input_size() or R2Plus1DClassifier((1, 21, 64, 64))


def test_loss_surface():
    f = FocalLoss(weight=torch.ones(2), gamma=2.0)
    l = LDAMLoss([300, 17000], max_m=0.5, weight=None, s=30)
    c = CELoss()
    assert (f.model_type, l.model_type, c.model_type) == ("Focal", "LDAM", "CE")
    np.testing.assert_allclose(l.m_list.numpy(), 0.5 * np.array([1.0, (300 / 17000) ** 0.25]), rtol=1e-6)
    for lf in (f, l, c):
        lf.update_weight(torch.tensor([0.3, 0.7]))
        assert torch.equal(lf.weight, torch.tensor([0.3, 0.7]))
    l.update_m_list([10, 1000])
    assert abs(l.m_list[0].item() - 0.5) < 1e-7
    with pytest.raises(AssertionError):
        FocalLoss(gamma=-1.0)


def test_imbalanced_sampler_rebalances():
    class DS:
        labels = [0] * 20 + [1] * 980

        def __len__(self):
            return len(self.labels)

    torch.manual_seed(0)
    s = ImbalancedDatasetSampler(DS())
    assert len(s) == 1000
    idx = list(iter(s))
    frac0 = sum(1 for i in idx if DS.labels[i] == 0) / len(idx)
    assert 0.42 < frac0 < 0.58          # 1/count weights => classes drawn about equally
    assert abs(s.weights[:20].sum().item() - 1.0) < 1e-9 and abs(s.weights[20:].sum().item() - 1.0) < 1e-9


def test_shards_and_windows():
    for n, w in ((1000, 8), (984, 4), (7, 8), (0, 2)):
        covered = []
        for r in range(w):
            covered += list(dpd.shard_range(n, r, w))
        assert covered == list(range(n))
    assert inference.num_windows(1024, 21, 3) == 1000
    assert inference.num_windows(10, 21, 3) == 0
    curve = inference.postprocess_curve([0.9, 0.9, 0.2, 0.7, 0.1], clip_len=2, frame_srt=1, fps=4)
    assert curve == [0, 0, 0, 0, 0.2, 0.7]                  # start-up suppression below fps frames
    m = R2Plus1DClassifier((3, 21, 128, 128), 2, [1, 2, 2, 1], False, 1.0)
    buckets = dpd.stage_buckets(m)
    assert len(buckets) == 6
    assert sum(p.numel() for b in buckets for p in b) == 1_587_523
    assert buckets[0][0] is m.linear[0].weight               # head first: its gradient is ready first


def test_multimodal_surface_cpu():
    """Config-4 model constructs on CPU with the reference's fusion-head structure; the 0D branch is
    plain PyTorch and runs on CPU, the video branch refuses to (no CPU fallback)."""
    from dp_b200.MultiModal import GradientBlending, MultiModalR2Plus1D, MultiModalR2Plus1D_GB, Transformer
    args_v = {"layer_sizes": [1, 1, 1, 1], "alpha": 0.01}
    args_t = dict(n_features=18, kernel_size=5, feature_dims=128, max_len=21, n_layers=1, n_heads=8,
                  dim_feedforward=256, dropout=0.0)
    m = MultiModalR2Plus1D(2, args_v, args_t)
    assert m.connector[0].in_features == 256 and m.classifier[-1].out_features == 2      # MultiModal.py:21-31
    g = MultiModalR2Plus1D_GB(2, args_v, args_t)
    assert g.use_stream == "multi-GB" and g.vis_model.linear[0].in_features == 128
    t = Transformer(n_features=18, feature_dims=128, max_len=21, n_heads=8, dim_feedforward=256, dropout=0.0).eval()
    out = t(torch.randn(3, 21, 18))
    assert out.shape == (3, 2) and torch.isfinite(out).all()
    assert t.encoder.src_mask.shape == (21, 21) and torch.isinf(t.encoder.src_mask[0, 1]) and t.encoder.src_mask[1, 0] == 0
    gb = GradientBlending(torch.nn.CrossEntropyLoss(), torch.nn.CrossEntropyLoss(), torch.nn.CrossEntropyLoss(), 0.1, 0.4, 0.5)
    lo = torch.randn(4, 2)
    y = torch.tensor([0, 1, 1, 0])
    ce = torch.nn.functional.cross_entropy(lo, y)
    assert abs(gb(lo, lo, lo, y).item() - ce.item()) < 1e-6                               # weights sum to 1
    if not torch.cuda.is_available():
        with pytest.raises(_lib.DpError):
            m(torch.zeros(2, 3, 9, 64, 64), torch.zeros(2, 21, 18))


def test_slowfast_surface_cpu(golden_dir):
    """Config-3 model: seed-42 construction gives the reference's 339 state-dict keys, shapes and weights
    (fixture from the unmodified reference); BatchNorm buffers carry the deterministic part of the reference's
    constructor probe."""
    from dp_b200.slowfast import Bottleneck3D, SlowFast
    gold = np.load(os.path.join(golden_dir, "slowfast_step.npz"))
    torch.manual_seed(42)
    m = SlowFast((3, 20, 64, 64), Bottleneck3D, [1, 2, 2, 1], 4, 1, 2, 1.0)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in gold["keys"]] and len(sd) == 339
    assert sum(p.numel() for p in m.parameters()) == 1_076_526
    for k, shape, ref in zip(gold["keys"], gold["shapes"], gold["summary"]):
        t = sd[str(k)]
        assert str(tuple(t.shape)) == str(shape), k
        if "running_" in str(k) or "num_batches" in str(k):
            continue
        f = t.detach().double().reshape(-1)
        assert abs(f.sum().item() - ref[0]) <= 1e-6 * max(1.0, ref[1]) and abs(f.abs().sum().item() - ref[1]) <= 1e-6 * max(1.0, ref[1]), k
    assert int(sd["encoder.slownet.layer0.1.num_batches_tracked"]) == 1
    assert torch.allclose(sd["encoder.fastnet.layer1.0.bn1.running_var"], torch.full((4,), 0.9))
    assert torch.allclose(sd["encoder.slownet.layer0.1.running_mean"], 0.1 * sd["encoder.slownet.layer0.0.bias"])
    assert m.classifier.classifier[0].in_features == 640


def test_every_layer_of_the_benchmark_has_a_tcgen05_plan():
    """Planner regression guard (runs without a GPU: the plan only needs the SM count): at the benchmark's shapes
    (B = 64, (3,21,128,128), [1,2,2,1]) every convolution must be served by the tcgen05 family in forward, data
    gradient and weight gradient -- a planner change that silently dropped a layer to the CUDA-core kernels cost 2x
    once -- and the stem must take its packed-rows fast path."""
    import ctypes as C
    from dp_b200 import _lib as L
    from dp_b200 import functional as Fn
    from oracle import r2plus1d_port as port

    lib = L.load()

    def out_dim(n, k, s, p):
        return (n + 2 * p - k) // s + 1

    stem, blocks = port.encoder_plan([1, 2, 2, 1], 1.0)
    seq = []

    def run_seq(layers, inp):
        for (name, cin, cout, k, s, p, slope) in layers:
            seq.append((name, cin, cout, k, s, p, inp))
            inp = tuple(out_dim(inp[i], k[i], s[i], p[i]) for i in range(3))
        return inp

    cur = run_seq(stem, (21, 128, 128))
    for b in blocks:
        o = run_seq(b["conv1"], cur)
        o = run_seq(b["conv2"], o)
        if b["shortcut"]:
            run_seq(b["shortcut"], cur)
        cur = o
    assert len(seq) == 32
    B = 64
    missing = []
    for (name, cin, cout, k, s, p, (T, H, W)) in seq:
        To, Ho, Wo = (out_dim(n, k[i], s[i], p[i]) for i, n in enumerate((T, H, W)))
        d = L.ConvDesc(B, T, H, W, cin, Fn.ceil16(cin), To, Ho, Wo, cout, Fn.ceil16(cout), *k, *s, *p, L.DP_BF16)
        if cin == 3:
            if not lib.dp_stem_supported(C.byref(d)):
                missing.append((name, "stem fast path"))
            continue
        for op, what in enumerate(("fwd", "dgrad", "wgrad")):
            if not lib.dp_conv_supported(C.byref(d), op, L.IMPL_TC):
                missing.append((name, what))
    assert not missing, missing


def test_every_slowfast_conv_has_a_tcgen05_plan(golden_dir):
    """All 74 convolutions of SlowFast [1,2,2,1] at the config-3 clip (geometries recorded from the reference by
    oracle/make_golden_r2.py) have tcgen05 plans for forward, data gradient and weight gradient at batch 2 and 64: the
    bf16 path never drops to the CUDA-core family.  Host-side planner only: runs on the CPU."""
    import ctypes as C
    import json
    from dp_b200 import _lib as L, functional as Fn
    lib = L.load()
    geoms = json.load(open(os.path.join(golden_dir, "slowfast_conv_geoms.json")))
    assert len(geoms) == 74
    for B in (2, 64):
        for g in geoms:
            cg = Fn.ConvGeom(g["C"], g["K"], tuple(g["kernel"]), tuple(g["stride"]), tuple(g["padding"]), B, g["T"], g["H"], g["W"],
                             L.DP_BF16)
            for op in range(3):
                assert lib.dp_conv_supported(C.byref(cg.desc), op, L.IMPL_TC), (B, g, op)


def test_plans_of_the_heavy_layers_keep_their_pipeline_shape():
    """dp_conv_describe_plan (no device needed): the five layers that carry 60 % of the step keep the pipeline features
    their measured speed depends on -- 256-pixel tiles where they were faster, two MMA-issuing
    warps, resident weights, statistics in epilogue registers, the narrow tail block of the 80-channel operands."""
    import ctypes as C
    from dp_b200 import _lib as L
    from dp_b200 import functional as Fn

    lib = L.load()

    def plan(cin, cout, k, s, p, thw, op, stats):
        T, H, W = thw
        o = [(n + 2 * p[i] - k[i]) // s[i] + 1 for i, n in enumerate((T, H, W))]
        d = L.ConvDesc(64, T, H, W, cin, Fn.ceil16(cin), *o, cout, Fn.ceil16(cout), *k, *s, *p, L.DP_BF16)
        buf = C.create_string_buffer(512)
        assert lib.dp_conv_describe_plan(C.byref(d), op, stats, buf, 512) == 0, (cin, cout, op)
        return dict(kv.split("=") for kv in buf.value.decode().split() if "=" in kv)

    full = (21, 64, 64)
    t = plan(45, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), full, 0, 1)            # stem temporal forward
    assert (t["MT"], t["dual"], t["reg_stats"], t["resident"], t["nloads"]) == ("2", "1", "1", "1", "1")
    s = plan(32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), full, 0, 1)            # conv2 spatial forward
    assert (s["Ntile"], s["reg_stats"], s["resident"], s["nloads"], s["nsub"]) == ("80", "1", "1", "3", "3")
    assert (s["MT"], s["dual"], s["st_bufs"]) == ("2", "1", "2")            # two issuers AND a slot per 128-row sub-tile
    g = plan(32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), full, 1, 0)            # conv2 spatial data gradient (80 -> 32)
    assert (g["CBt"], g["dual"], g["acc_bufs"], g["resident"]) == ("16", "1", "4", "1")
    u = plan(72, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), full, 1, 0)            # conv2 temporal data gradient (32 -> 80)
    assert (u["MT"], u["dual"], u["st_bufs"]) == ("2", "1", "4")
    for v in (t, s, g, u):
        assert int(v["smem"]) <= 232448 and int(v["tmem_cols"]) <= 512 and v["grid"] == "148"
    # the structure behind dp_bn_fin: 4 int32, a double, 2 pointers, 2 floats, 10 pointers
    assert C.sizeof(L.BnFin) == 16 + 8 + 16 + 8 + 80


def test_forward_channel_split_is_an_option_and_off_by_default():
    """dp_conv_describe_plan (no device needed) reports the output-channel split of the forward conv: off by default
    (measured neutral on the whole step); with `tc_nsplit` = 1 only the 64 -> 144 spatial convs whose
    166 KB of weights do not fit beside the pipeline AND that have >= 4 tiles per CTA split (80 + 64 channels, both halves
    resident on the same grid); layers with resident weights, or with several N tiles, never do."""
    import ctypes as C
    from dp_b200 import _lib as L
    from dp_b200 import functional as Fn

    lib = L.load()

    def split_of(B, cin, cout, k, p, thw):
        T, H, W = thw
        o = [(n + 2 * p[i] - k[i]) + 1 for i, n in enumerate((T, H, W))]
        d = L.ConvDesc(B, T, H, W, cin, Fn.ceil16(cin), *o, cout, Fn.ceil16(cout), *k, 1, 1, 1, *p, L.DP_BF16)
        buf = C.create_string_buffer(1024)
        assert lib.dp_conv_describe_plan(C.byref(d), 0, 1, buf, 1024) == 0, (cin, cout)
        return int(dict(kv.split("=") for kv in buf.value.decode().split() if "=" in kv)["split"])

    spatial, temporal = ((1, 3, 3), (0, 1, 1)), ((3, 1, 1), (1, 0, 0))
    default = L.get_option("tc_nsplit")
    try:
        assert default == 0
        assert split_of(64, 64, 144, *spatial, (11, 32, 32)) == 0
        L.set_option("tc_nsplit", 1)
        assert split_of(64, 64, 144, *spatial, (11, 32, 32)) == 80       # conv3 spatial: 5632 tiles
        assert split_of(64, 64, 144, *spatial, (6, 16, 16)) == 80        # conv4 spatial: 768 tiles
        assert split_of(2, 64, 144, *spatial, (11, 32, 32)) == 0         # 176 tiles: not worth a second launch
        assert split_of(64, 32, 72, *spatial, (21, 64, 64)) == 0         # weights resident already
        assert split_of(64, 144, 64, *temporal, (11, 32, 32)) == 0
        assert split_of(64, 128, 288, *spatial, (3, 8, 8)) == 0          # two N tiles: not covered
        L.set_option("tc_nsplit", 2)
        assert split_of(2, 64, 144, *spatial, (11, 32, 32)) == 80        # 2: whenever the halves become resident (tests)
    finally:
        L.set_option("tc_nsplit", default)
