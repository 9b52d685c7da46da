"""Kernel-level parity on the GPU, every call through the C ABI (ctypes): each CUDA kernel against a
plain PyTorch fp32 restatement of the same op on the same seeded inputs.

Tolerances: fp32 validation kernels 1e-4 relative (north_star); bf16 kernels compare against the fp32
op evaluated on the SAME bf16-rounded operands, so only accumulation order and the final bf16 rounding
of the output differ: 2^-8 relative to the tensor's scale."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

import dp_b200
from dp_b200 import _lib as L
from dp_b200 import functional as Fn

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def to_int(x, dtype):
    """NCDHW fp32 -> internal via the kernel."""
    return Fn._to_internal_raw(x.contiguous(), dtype)


def from_int(x, c):
    return Fn._to_ncdhw_raw(x, c)


# ---- R(2+1)D [1,2,2,1] layer geometries at T=21,128^2 scaled to a small clip ------------
# (C, K, kernel, stride, padding, T, H, W)
GEOMS = [
    (3, 45, (1, 7, 7), (1, 2, 2), (0, 3, 3), 5, 32, 32),      # stem spatial
    (45, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), 5, 16, 16),     # stem temporal
    (32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), 5, 16, 16),     # conv2 spatial
    (72, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), 5, 16, 16),     # conv2 temporal
    (32, 115, (1, 3, 3), (1, 2, 2), (0, 1, 1), 5, 16, 16),    # conv3 down spatial
    (115, 64, (3, 1, 1), (2, 1, 1), (1, 0, 0), 5, 8, 8),      # conv3 down temporal
    (64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), 3, 8, 8),      # conv3/4 spatial
    (144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 3, 8, 8),      # conv3/4 temporal
    (32, 21, (1, 1, 1), (1, 2, 2), (0, 0, 0), 5, 16, 16),     # shortcut spatial
    (21, 64, (1, 1, 1), (2, 1, 1), (0, 0, 0), 5, 8, 8),       # shortcut temporal
    (64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1), 3, 8, 8),      # conv5 down spatial
    (230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0), 3, 4, 4),     # conv5 down temporal
    (128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 4, 4),     # conv5 spatial (N > 256)
    (288, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 4, 4),     # conv5 temporal
    (320, 512, (1, 1, 1), (1, 2, 2), (0, 0, 0), 5, 8, 8),     # SlowFast layer4 shortcut: both channel counts > 256 (wgrad in K halves)
]
# full-resolution shapes for the tensor-core family (tile edges, halo boxes, 64-wide rows)
GEOMS_BIG = [
    (32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), 21, 64, 64),
    (72, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), 21, 64, 64),
    (45, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), 21, 64, 64),
    (64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), 11, 32, 32),
    (144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 11, 32, 32),
    (64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), 6, 16, 16),
    (128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), 3, 8, 8),
    (288, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0), 3, 8, 8),
    (3, 45, (1, 7, 7), (1, 2, 2), (0, 3, 3), 21, 128, 128),
    (32, 115, (1, 3, 3), (1, 2, 2), (0, 1, 1), 21, 64, 64),
    (115, 64, (3, 1, 1), (2, 1, 1), (1, 0, 0), 21, 32, 32),
    (32, 21, (1, 1, 1), (1, 2, 2), (0, 0, 0), 21, 64, 64),
    (21, 64, (1, 1, 1), (2, 1, 1), (0, 0, 0), 21, 32, 32),
]


def make_case(geom, B, dtype, seed=0):
    Cc, K, k, s, p, T, H, W = geom
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cc, T, H, W, generator=g)
    w = torch.randn(K, Cc, *k, generator=g) * math.sqrt(2.0 / (Cc * k[0] * k[1] * k[2]))
    if dtype == torch.bfloat16:  # compare on identical (bf16-representable) operands
        x = x.bfloat16().float()
        w = w.bfloat16().float()
    return x.to(DEV), w.to(DEV)


def run_conv(geom, B, dtype, impl, x, w, dy=None, addend=None):
    """Returns y (NCDHW fp32) and, if dy given, (dx, dw)."""
    Cc, K, k, s, p, T, H, W = geom
    lib = L.load()
    xi = to_int(x, dtype)
    gm = Fn.conv_geom(Cc, K, k, s, p, xi)
    d = gm.desc
    wf = torch.empty((d.Kp, gm.taps, d.Cp), dtype=dtype, device=DEV)
    wd = torch.empty((d.Cp, gm.taps, d.Kp), dtype=dtype, device=DEV)
    L.check(lib.dp_pack_weights(C.byref(d), w.contiguous().data_ptr(), wf.data_ptr(), wd.data_ptr(), L.stream_ptr()), "pack")
    y = torch.empty(gm.out_shape, dtype=dtype, device=DEV)
    part = torch.zeros((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=DEV)
    nparts = C.c_int(0)
    L.check(lib.dp_conv_fwd(C.byref(d), xi.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(), C.byref(nparts),
                            impl, L.stream_ptr()), "fwd")
    out = {"y": from_int(y, K), "y_int": y, "part": part[:nparts.value].clone(), "rows": gm.rows_out, "K": K}
    if dy is not None:
        dyi = to_int(dy, dtype)
        dx = torch.empty_like(xi)
        ad = to_int(addend, dtype) if addend is not None else None
        L.check(lib.dp_conv_dgrad(C.byref(d), dyi.data_ptr(), wd.data_ptr(), None if ad is None else ad.data_ptr(),
                                  dx.data_ptr(), impl, L.stream_ptr()), "dgrad")
        dw = torch.empty_like(w)
        nbytes = int(lib.dp_conv_wgrad_workspace(C.byref(d), impl))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=DEV)
        L.check(lib.dp_conv_wgrad(C.byref(d), xi.data_ptr(), dyi.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(),
                                  impl, L.stream_ptr()), "wgrad")
        out["dx"] = from_int(dx, Cc)
        out["dw"] = dw
    return out


def torch_ref(geom, x, w, dy=None):
    Cc, K, k, s, p, T, H, W = geom
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        xr = x.clone().double().requires_grad_(True)   # fp64: an exact-enough reference for both modes
        wr = w.clone().double().requires_grad_(True)
        y = F.conv3d(xr, wr, None, s, p)
        out = {"y": y.detach()}
        if dy is not None:
            y.backward(dy.double())
            out["dx"], out["dw"] = xr.grad, wr.grad
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out


@pytest.mark.parametrize("geom", GEOMS, ids=lambda g: f"{g[0]}to{g[1]}_k{g[2]}_s{g[3]}")
def test_conv_simt_fp32(geom):
    B = 2
    x, w = make_case(geom, B, torch.float32)
    ref = torch_ref(geom, x, w)
    dy = torch.randn_like(ref["y"], dtype=torch.float32)
    ref = torch_ref(geom, x, w, dy)
    got = run_conv(geom, B, torch.float32, L.IMPL_SIMT, x, w, dy)
    assert rel_err(got["y"], ref["y"]) < 1e-5
    assert rel_err(got["dx"], ref["dx"]) < 1e-5
    assert rel_err(got["dw"], ref["dw"]) < 1e-5
    # BN partial statistics of the conv output
    s = got["part"].double().sum(0)
    yy = ref["y"].double()
    assert rel_err(s[0, :geom[1]], yy.sum(dim=(0, 2, 3, 4))) < 1e-4
    assert rel_err(s[1, :geom[1]], (yy * yy).sum(dim=(0, 2, 3, 4))) < 1e-4


@pytest.mark.parametrize("geom", GEOMS, ids=lambda g: f"{g[0]}to{g[1]}_k{g[2]}_s{g[3]}")
def test_conv_simt_bf16(geom):
    B = 2
    x, w = make_case(geom, B, torch.bfloat16)
    ref = torch_ref(geom, x, w)
    dy = torch.randn_like(ref["y"], dtype=torch.float32).bfloat16().float()
    ref = torch_ref(geom, x, w, dy)
    got = run_conv(geom, B, torch.bfloat16, L.IMPL_SIMT, x, w, dy)
    assert rel_err(got["y"], ref["y"]) < 2 ** -7
    assert rel_err(got["dx"], ref["dx"]) < 2 ** -7
    assert rel_err(got["dw"], ref["dw"]) < 1e-4   # fp32 output, fp32 accumulation


def _tc_check(geom, B, seed=0):
    lib = L.load()
    x, w = make_case(geom, B, torch.bfloat16, seed)
    ref = torch_ref(geom, x, w)
    dy = torch.randn_like(ref["y"], dtype=torch.float32).bfloat16().float()
    ref = torch_ref(geom, x, w, dy)
    Cc, K, k, s, p, T, H, W = geom
    xi = to_int(x, torch.bfloat16)
    gm = Fn.conv_geom(Cc, K, k, s, p, xi)
    d = gm.desc
    res = {}
    if lib.dp_conv_supported(C.byref(d), 0, L.IMPL_TC):
        wf, wd = Fn.pack_weights(w.contiguous(), gm, torch.bfloat16, None)
        y = torch.full(gm.out_shape, float("nan"), dtype=torch.bfloat16, device=DEV)
        part = torch.zeros((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=DEV)
        nparts = C.c_int(0)
        L.check(lib.dp_conv_fwd(C.byref(d), xi.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(),
                                C.byref(nparts), L.IMPL_TC, L.stream_ptr()), "tc fwd")
        torch.cuda.synchronize()
        yo = from_int(y, K)
        res["fwd"] = rel_err(yo, ref["y"])
        assert torch.isfinite(y.float()).all(), "tc fwd left unwritten (NaN) output elements"
        assert (y[..., K:] == 0).all(), "padded output channels must be exactly zero"
        st = part[:nparts.value].double().sum(0)
        # statistics come from the fp32 accumulators (register mode) or from the bf16-rounded tile (tensor-core / CUDA-core
        # modes): they must match the fp64 convolution, or the rounded output, on the scale of sum|y| (zero-mean sums cancel)
        def stat_err(yy):
            yy = yy.double()
            e_sum = float((st[0, :K] - yy.sum(dim=(0, 2, 3, 4))).abs().max() / yy.abs().sum(dim=(0, 2, 3, 4)).max())
            return e_sum, rel_err(st[1, :K], (yy * yy).sum(dim=(0, 2, 3, 4)))
        res["sum"], res["sq"] = min(stat_err(ref["y"]), stat_err(yo), key=lambda e: e[0] + e[1])
        assert res["fwd"] < 2 ** -7, res
        assert res["sum"] < 1e-3 and res["sq"] < 1e-3, res
    if lib.dp_conv_supported(C.byref(d), 1, L.IMPL_TC):
        # one launch per stride-parity class (dp_conv_dgrad, [Cp][taps][Kp] weights) and -- strided geometries -- every class
        # in ONE launch over class-packed weights (dp_conv_dgrad_classes): both against the fp64 conv_transpose
        wd = torch.empty((d.Cp, gm.taps, d.Kp), dtype=torch.bfloat16, device=DEV)
        L.check(lib.dp_pack_weights(C.byref(d), w.contiguous().data_ptr(), None, wd.data_ptr(), L.stream_ptr()), "pack")
        n_cls = int(lib.dp_dgrad_classes_weight_elems(C.byref(d), L.IMPL_TC))
        if n_cls:
            wc = torch.empty(n_cls, dtype=torch.bfloat16, device=DEV)
            L.check(lib.dp_pack_weights_dgrad_classes(C.byref(d), w.contiguous().data_ptr(), wc.data_ptr(), L.stream_ptr()), "pack classes")
        dyi = to_int(dy, torch.bfloat16)
        for use_add in (False, True):
            ad = torch.randn_like(x).bfloat16().float() if use_add else None
            adi = to_int(ad, torch.bfloat16) if use_add else None
            want = ref["dx"] + (ad.double() if use_add else 0)
            for path in (("per_class", "classes") if n_cls else ("per_class",)):
                dx = torch.full(gm.in_shape, float("nan"), dtype=torch.bfloat16, device=DEV)
                if path == "classes":
                    L.check(lib.dp_conv_dgrad_classes(C.byref(d), dyi.data_ptr(), wc.data_ptr(),
                                                      None if adi is None else adi.data_ptr(), dx.data_ptr(), L.stream_ptr()),
                            "tc dgrad classes")
                else:
                    L.check(lib.dp_conv_dgrad(C.byref(d), dyi.data_ptr(), wd.data_ptr(),
                                              None if adi is None else adi.data_ptr(), dx.data_ptr(), L.IMPL_TC,
                                              L.stream_ptr()), "tc dgrad")
                torch.cuda.synchronize()
                key = ("dgrad_add" if use_add else "dgrad") + ("_cls" if path == "classes" else "")
                assert torch.isfinite(dx.float()).all(), f"{key}: unwritten (NaN) elements of dx"
                assert (dx[..., Cc:] == 0).all() or use_add, f"{key}: padded input channels must stay zero"
                res[key] = rel_err(from_int(dx, Cc), want)
                assert res[key] < 2 ** -7, res
        if (d.st, d.sh, d.sw) != (1, 1, 1) and (T % d.st == 0 or True):
            res["classes_in_one_launch"] = bool(n_cls)
    if lib.dp_conv_supported(C.byref(d), 2, L.IMPL_TC):
        dyi = to_int(dy, torch.bfloat16)
        dw = torch.empty_like(w)
        nbytes = int(lib.dp_conv_wgrad_workspace(C.byref(d), L.IMPL_TC))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=DEV)
        L.check(lib.dp_conv_wgrad(C.byref(d), xi.data_ptr(), dyi.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(),
                                  L.IMPL_TC, L.stream_ptr()), "tc wgrad")
        torch.cuda.synchronize()
        res["wgrad"] = rel_err(dw, ref["dw"])
        assert res["wgrad"] < 1e-3, res
    return res


@pytest.mark.parametrize("geom", GEOMS + GEOMS_BIG, ids=lambda g: f"{g[0]}to{g[1]}_k{g[2]}_s{g[3]}_{g[5]}x{g[6]}")
def test_conv_tcgen05(geom):
    res = _tc_check(geom, B=2)
    print(geom, res)
    assert "fwd" in res and "wgrad" in res, "every geometry of the path has a tcgen05 forward and weight-gradient kernel"
    if geom[3] != (1, 1, 1) and geom[0] <= 256 and geom[2] != (1, 1, 1):   # (1x1x1 shortcuts: one class + a copy, by design)
        assert res.get("classes_in_one_launch"), "strided data gradients of the R(2+1)D path run as one launch"


def test_conv_tcgen05_modes():
    """Same geometry through the per-tap path and the halo-reuse path must agree with the reference."""
    try:
        for halo in (0, 1):
            L.set_option("tc_halo", halo)
            L.set_option("wg_halo", halo)
            for geom in (GEOMS_BIG[0], GEOMS_BIG[1], GEOMS_BIG[3]):
                _tc_check(geom, B=1, seed=3)
    finally:
        L.set_option("tc_halo", 1)
        L.set_option("wg_halo", 1)


@pytest.mark.parametrize("shape", [(2, 5, 32, 32), (2, 21, 128, 128), (1, 3, 50, 38)])
@pytest.mark.parametrize("u8", [False, True])
def test_stem_fast_path(shape, u8):
    """Packed-rows stem (csrc/stem.cu): forward + BN partials + weight gradient against fp64 conv3d on the same
    bf16-rounded operands, from NCDHW fp32 clips and from uint8 frames."""
    B, T, H, W = shape
    lib = L.load()
    k, s_, p_ = (1, 7, 7), (1, 2, 2), (0, 3, 3)
    g = torch.Generator().manual_seed(4)
    frames = torch.randint(0, 256, (B, T, H, W, 3), generator=g, dtype=torch.uint8)
    mean = (90.0, 98.0, 102.0)
    x = (frames.float() - torch.tensor(mean)).permute(0, 4, 1, 2, 3).contiguous()     # NCDHW, exactly bf16-representable
    w = (torch.randn(45, 3, *k, generator=g) * math.sqrt(2.0 / 147)).bfloat16().float()
    x, w, frames = x.to(DEV), w.to(DEV), frames.to(DEV)
    cfg = Fn.LayerCfg(3, 45, k, s_, p_, 1.0)
    with dp_b200.compute_mode("bf16", "auto"):
        gm = Fn.stem_geom(cfg, B, T, H, W)
    assert gm is not None, "stem fast path must cover the R(2+1)D stem geometry"
    d = gm.desc
    xp = Fn.stem_pack_input(frames if u8 else x, gm, mean)
    wv, _ = Fn.pack_weights(w, gm, torch.bfloat16, None)
    y = torch.full(gm.out_shape, float("nan"), dtype=torch.bfloat16, device=DEV)
    part = torch.zeros((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=DEV)
    nparts = C.c_int(0)
    L.check(lib.dp_stem_conv_fwd(C.byref(d), xp.data_ptr(), wv.data_ptr(), y.data_ptr(), part.data_ptr(),
                                 C.byref(nparts), L.stream_ptr()), "stem fwd")
    geom = (3, 45, k, s_, p_, T, H, W)
    ref = torch_ref(geom, x, w)
    yo = from_int(y, 45)
    assert torch.isfinite(y.float()).all()
    assert rel_err(yo, ref["y"]) < 2 ** -7
    assert (y[..., 45:] == 0).all()
    st = part[:nparts.value].double().sum(0)
    assert rel_err(st[0, :45], yo.double().sum(dim=(0, 2, 3, 4))) < 1e-3
    assert rel_err(st[1, :45], (yo.double() ** 2).sum(dim=(0, 2, 3, 4))) < 1e-3
    dy = torch.randn_like(ref["y"], dtype=torch.float32).bfloat16().float()
    ref = torch_ref(geom, x, w, dy)
    dyi = to_int(dy, torch.bfloat16)
    dw = torch.full_like(w, float("nan"))
    ws = torch.empty(int(lib.dp_stem_wgrad_workspace(C.byref(d))), dtype=torch.uint8, device=DEV)
    L.check(lib.dp_stem_conv_wgrad(C.byref(d), xp.data_ptr(), dyi.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(),
                                   L.stream_ptr()), "stem wgrad")
    assert rel_err(dw, ref["dw"]) < 1e-3


# ---- BatchNorm / activation ------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C_", [32, 45, 72, 288])
def test_bn_act_fwd_bwd(dtype, C_):
    lib = L.load()
    g = torch.Generator().manual_seed(1)
    B, T, H, W = 2, 3, 8, 8
    slope, slope_res = 0.01, 0.3
    y = (torch.randn(B, C_, T, H, W, generator=g) * 2 + 0.5)
    res = torch.randn(B, C_, T, H, W, generator=g)
    dz = torch.randn(B, C_, T, H, W, generator=g)
    gamma = torch.rand(C_, generator=g) + 0.5
    beta = torch.randn(C_, generator=g) * 0.1
    if dtype == torch.bfloat16:
        y, res, dz = y.bfloat16().float(), res.bfloat16().float(), dz.bfloat16().float()
    y, res, dz, gamma, beta = (t.to(DEV) for t in (y, res, dz, gamma, beta))
    for use_res in (False, True):
        # reference
        yr = y.clone().requires_grad_(True)
        rr = res.clone().requires_grad_(True)
        gr = gamma.clone().requires_grad_(True)
        br = beta.clone().requires_grad_(True)
        rm, rv = torch.zeros(C_, device=DEV), torch.ones(C_, device=DEV)
        z = F.leaky_relu(F.batch_norm(yr, rm, rv, gr, br, True, 0.1, 1e-5), slope)
        if use_res:
            z = F.leaky_relu(z + rr, slope_res)
        z.backward(dz)
        # ours
        Cp = Fn.ceil16(C_)
        yi, ri, dzi = to_int(y, dtype), to_int(res, dtype), to_int(dz, dtype)
        rows = B * T * H * W
        part = torch.zeros((L.DP_MAX_PARTS, 2, Cp), dtype=torch.float32, device=DEV)
        nparts = C.c_int(0)
        st_ = L.stream_ptr()
        L.check(lib.dp_bn_stats(yi.data_ptr(), rows, Cp, Fn._code(yi), part.data_ptr(), C.byref(nparts), st_))
        stats = torch.zeros((4, Cp), dtype=torch.float32, device=DEV)
        rm2, rv2 = torch.zeros(C_, device=DEV), torch.ones(C_, device=DEV)
        L.check(lib.dp_bn_finalize(part.data_ptr(), nparts.value, C_, Cp, float(rows), gamma.data_ptr(), beta.data_ptr(),
                                   1e-5, 0.1, rm2.data_ptr(), rv2.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                                   stats[2].data_ptr(), stats[3].data_ptr(), st_))
        zi = torch.empty_like(yi)
        L.check(lib.dp_bn_act_apply(yi.data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), slope,
                                    ri.data_ptr() if use_res else None, slope_res, zi.data_ptr(), rows, Cp,
                                    Fn._code(yi), st_))
        tol = 1e-4 if dtype == torch.float32 else 2 ** -7
        assert rel_err(from_int(zi, C_), z.detach()) < tol
        assert rel_err(rm2, rm) < 1e-5 and rel_err(rv2, rv) < 1e-5
        assert (zi[..., C_:] == 0).all()
        L.check(lib.dp_bn_act_bwd_reduce(dzi.data_ptr(), yi.data_ptr(), zi.data_ptr() if use_res else None,
                                         stats[2].data_ptr(), stats[3].data_ptr(), stats[0].data_ptr(),
                                         stats[1].data_ptr(), slope, slope_res, part.data_ptr(), C.byref(nparts), rows,
                                         Cp, Fn._code(yi), st_))
        dgamma, dbeta = torch.empty(C_, device=DEV), torch.empty(C_, device=DEV)
        coef = torch.empty((2, Cp), device=DEV)
        L.check(lib.dp_bn_bwd_finalize(part.data_ptr(), nparts.value, C_, Cp, float(rows), stats[0].data_ptr(),
                                       stats[1].data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), st_))
        dyi = torch.empty_like(yi)
        dres = torch.empty_like(yi) if use_res else None
        L.check(lib.dp_bn_act_bwd_apply(dzi.data_ptr(), yi.data_ptr(), zi.data_ptr() if use_res else None,
                                        stats[2].data_ptr(), stats[3].data_ptr(), stats[0].data_ptr(),
                                        stats[1].data_ptr(), coef.data_ptr(), slope, slope_res, dyi.data_ptr(),
                                        None if dres is None else dres.data_ptr(), rows, Cp, Fn._code(yi), st_))
        gtol = 2e-4 if dtype == torch.float32 else 2 ** -6
        assert rel_err(dgamma, gr.grad) < gtol
        assert rel_err(dbeta, br.grad) < gtol
        assert rel_err(from_int(dyi, C_), yr.grad) < gtol
        if use_res:
            assert rel_err(from_int(dres, C_), rr.grad) < gtol


def test_layout_roundtrip_and_u8():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 45, 3, 6, 10, generator=g).to(DEV)
    for dtype in (torch.float32, torch.bfloat16):
        xi = to_int(x, dtype)
        assert xi.shape == (2, 3, 6, 10, 48)
        want = x.permute(0, 2, 3, 4, 1).to(dtype)
        assert torch.equal(xi[..., :45], want)
        assert (xi[..., 45:] == 0).all()
        back = from_int(xi, 45)
        assert torch.equal(back, want.permute(0, 4, 1, 2, 3).float())
    frames = torch.randint(0, 256, (2, 3, 8, 8, 3), generator=g, dtype=torch.uint8).to(DEV)
    with dp_b200.compute_mode("bf16"):
        xi = Fn.frames_u8_to_internal(frames, (90.0, 98.0, 102.0))
    want = (frames.float() - torch.tensor([90.0, 98.0, 102.0], device=DEV)).bfloat16()
    assert torch.equal(xi[..., :3], want) and (xi[..., 3:] == 0).all()


def test_avgpool():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 128, 3, 8, 8, generator=g).to(DEV)
    for dtype in (torch.float32, torch.bfloat16):
        xi = Fn.tag(to_int(x, dtype), 128).requires_grad_(True)
        out = Fn.AvgPoolFn.apply(xi, 128)
        want = xi.detach().float()[..., :128].mean(dim=(1, 2, 3))
        assert rel_err(out, want) < 1e-5
        go = torch.randn(3, 128, generator=g).to(DEV)
        out.backward(go)
        wantg = (go / 192.0)[:, None, None, None, :].expand(3, 3, 8, 8, 128)
        assert rel_err(xi.grad.float()[..., :128], wantg.to(dtype).float()) < 1e-6


def test_fused_clip_adamw_matches_torch():
    from dp_b200.optim import FusedClipAdamW
    g = torch.Generator().manual_seed(4)
    shapes = [(45, 3, 1, 7, 7), (45,), (32, 45, 3, 1, 1), (7,), (64, 128)]
    ps_a = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    oa = FusedClipAdamW(ps_a, lr=2e-4, max_norm=1.0)
    ob = torch.optim.AdamW(ps_b, lr=2e-4)
    for step in range(3):
        for pa, pb in zip(ps_a, ps_b):
            gr = torch.randn(pa.shape, generator=g).to(DEV) * (5.0 if step == 0 else 0.01)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        norm_b = torch.nn.utils.clip_grad_norm_(ps_b, 1.0)
        ob.step()
        oa.step()
        assert rel_err(oa.last_grad_norm[0], norm_b) < 1e-5
        for pa, pb in zip(ps_a, ps_b):
            assert rel_err(pa.detach(), pb.detach()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("geom", [(72, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), 9, 16, 16),
                                  (32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), 5, 16, 16),
                                  (45, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), 6, 16, 16),
                                  (64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), 4, 8, 8),
                                  (32, 115, (1, 3, 3), (1, 2, 2), (0, 1, 1), 4, 16, 16)],
                         ids=lambda g: f"{g[0]}to{g[1]}_k{g[2]}_s{g[3]}")
@pytest.mark.parametrize("slope", [1.0, 0.01])
@pytest.mark.parametrize("use_add", [False, True])
def test_dgrad_bnstats(geom, slope, use_add):
    """dp_conv_dgrad_bnstats == dp_conv_dgrad followed by the BN-backward reduction of the producing layer
    (fused in the tcgen05 epilogue on stride-1 geometries, composed otherwise)."""
    lib = L.load()
    B = 3
    x, w = make_case(geom, B, torch.bfloat16, 5)
    ref = torch_ref(geom, x, w)
    dy = torch.randn_like(ref["y"], dtype=torch.float32).bfloat16().float()
    Cc, K, k, s, p, T, H, W = geom
    xi = to_int(x, torch.bfloat16)
    gm = Fn.conv_geom(Cc, K, k, s, p, xi)
    d = gm.desc
    wd = torch.empty((d.Cp, gm.taps, d.Kp), dtype=torch.bfloat16, device=DEV)      # the [Cp][taps][Kp] layout this entry point takes
    L.check(lib.dp_pack_weights(C.byref(d), w.contiguous().data_ptr(), None, wd.data_ptr(), L.stream_ptr()), "pack")
    dyi = to_int(dy, torch.bfloat16)
    g = torch.Generator(device="cpu").manual_seed(11)
    yprev = torch.randn(xi.shape, generator=g).to(DEV).bfloat16()
    yprev[..., Cc:] = 0
    add = None
    if use_add:
        add = torch.randn(xi.shape, generator=g).to(DEV).bfloat16()
        add[..., Cc:] = 0
    ss = torch.zeros((2, d.Cp), dtype=torch.float32, device=DEV)
    ss[0, :Cc] = torch.rand(Cc, generator=g).to(DEV) + 0.5
    ss[1, :Cc] = torch.randn(Cc, generator=g).to(DEV) * 0.3
    dx = torch.full(xi.shape, float("nan"), dtype=torch.bfloat16, device=DEV)
    part = torch.zeros((L.DP_MAX_PARTS, 2, d.Cp), dtype=torch.float32, device=DEV)
    nparts = C.c_int(0)
    L.check(lib.dp_conv_dgrad_bnstats(C.byref(d), dyi.data_ptr(), wd.data_ptr(), None if add is None else add.data_ptr(),
                                      dx.data_ptr(), yprev.data_ptr(), ss.data_ptr(), slope, part.data_ptr(),
                                      C.byref(nparts), L.IMPL_AUTO, L.stream_ptr()), "dgrad_bnstats")
    dx2 = torch.empty_like(dx)
    L.check(lib.dp_conv_dgrad(C.byref(d), dyi.data_ptr(), wd.data_ptr(), None if add is None else add.data_ptr(),
                              dx2.data_ptr(), L.IMPL_AUTO, L.stream_ptr()), "dgrad")
    torch.cuda.synchronize()
    assert torch.equal(dx, dx2), "the fused epilogue must not change the data gradient"
    st = part[:nparts.value].double().sum(0)
    # fp64 reference on the exact (unrounded) gradient: conv_transpose of the bf16 operands (+ addend)
    gref = torch_ref(geom, x, w, dy)["dx"].double()
    gi = gref.permute(0, 2, 3, 4, 1)
    if add is not None:
        gi = gi + add[..., :Cc].double()
    yp = yprev[..., :Cc].double()
    u = yp * ss[0, :Cc].double() + ss[1, :Cc].double()
    gp = gi * torch.where(u > 0, 1.0, slope)
    s0, s1 = gp.sum(dim=(0, 1, 2, 3)), (gp * yp).sum(dim=(0, 1, 2, 3))
    scale0 = gp.abs().sum(dim=(0, 1, 2, 3)).max()
    scale1 = (gp * yp).abs().sum(dim=(0, 1, 2, 3)).max()
    e0 = float((st[0, :Cc] - s0).abs().max() / scale0)
    e1 = float((st[1, :Cc] - s1).abs().max() / scale1)
    assert e0 < 2e-3 and e1 < 2e-3, (e0, e1)
    assert float(st[:, Cc:].abs().max()) == 0.0 if d.Cp > Cc else True


@pytest.mark.parametrize("opts", [{"tc_mt": 1}, {"tc_mt": 2}, {"tc_dual_mma": 0}, {"tc_tail": 0}, {"tc_acc4": 0},
                                  {"tc_lps_max": 1}, {"tc_reg_stats": 0}, {"tc_st_bufs": 1}],
                         ids=lambda o: ",".join(f"{k}={v}" for k, v in o.items()))
def test_conv_tcgen05_plan_variants(opts):
    """Every pipeline shape the planner can pick (128- / 256-pixel tiles, one or two MMA-issuing warps, narrow tail block
    on or off, 2 or 4 accumulators, one load or a whole tile per stage, statistics in registers or on the tensor core,
    one or two staging buffers) must give the same results against the fp64 reference."""
    defaults = {k: L.get_option(k) for k in opts}
    try:
        for k, v in opts.items():
            L.set_option(k, v)
        for geom in (GEOMS_BIG[0], GEOMS_BIG[1], GEOMS_BIG[2], GEOMS_BIG[4], GEOMS_BIG[9]):
            _tc_check(geom, B=1, seed=7)
    finally:
        for k, v in defaults.items():
            L.set_option(k, v)


def test_conv_fwd_output_channel_split():
    """Forward conv whose weights do not fit beside the pipeline (64 -> 144 channels, 9 taps): the output channels go in two
    launches ([0, 80) and [80, 144)) with resident weights, into channel sub-ranges of the same destination and of the same
    statistics partials.  dp_conv_describe_plan reports the split; output, padded channels and statistics over all 144
    channels are checked against the fp64 convolution by the common checker (R2Plus1D.py:53, conv3/conv4 spatial convs)."""
    import ctypes as C
    lib = L.load()
    default = L.get_option("tc_nsplit")
    try:
        for B, geom in ((8, GEOMS_BIG[3]), (16, GEOMS_BIG[5])):
            Cc, K, k, s, p, T, H, W = geom
            L.set_option("tc_nsplit", 2)      # 2: split whenever it makes the weights resident (1 also asks for >= 4 tiles per CTA)
            o = [(n + 2 * p[i] - k[i]) // s[i] + 1 for i, n in enumerate((T, H, W))]
            d = L.ConvDesc(B, T, H, W, Cc, Fn.ceil16(Cc), *o, K, Fn.ceil16(K), *k, *s, *p, L.DP_BF16)
            buf = C.create_string_buffer(1024)
            assert lib.dp_conv_describe_plan(C.byref(d), 0, 1, buf, 1024) == 0
            assert " split=80" in buf.value.decode(), buf.value
            n0 = lib.dp_launch_count()
            res = _tc_check(geom, B=B, seed=11)
            assert res["fwd"] < 2 ** -7 and res["sum"] < 1e-3 and res["sq"] < 1e-3, res
            L.set_option("tc_nsplit", 0)
            n1 = lib.dp_launch_count()
            _tc_check(geom, B=B, seed=11)
            assert (n1 - n0) - (lib.dp_launch_count() - n1) == 1, "the split forward is exactly one launch more"
    finally:
        L.set_option("tc_nsplit", default)


def _fin_fwd(d, rows, gamma, beta, rm, rv, stats, ticket):
    return L.BnFin(kind=1, C=d.K, Cp=d.Kp, coef_zero=0, count=float(rows), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                   eps=1e-5, momentum=0.1, running_mean=rm.data_ptr(), running_var=rv.data_ptr(),
                   mean=stats[0].data_ptr(), rstd=stats[1].data_ptr(), scale=stats[2].data_ptr(), shift=stats[3].data_ptr(),
                   dgamma=None, dbeta=None, coef=None, ticket=ticket.data_ptr())


@pytest.mark.parametrize("geom", [(45, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), 6, 16, 16),     # register statistics, one chunk pair
                                  (32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), 21, 64, 64),    # ... 80 channels, 148 CTAs, 256-pixel tiles
                                  (64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), 4, 8, 8),      # CUDA-core statistics from the staged tile
                                  (32, 115, (1, 3, 3), (1, 2, 2), (0, 1, 1), 4, 16, 16),    # statistics MMAs (Gram + ones)
                                  (128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 4, 4)],    # two N tiles
                         ids=lambda g: f"{g[0]}to{g[1]}_k{g[2]}_s{g[3]}_{g[5]}x{g[6]}")
@pytest.mark.parametrize("impl", ["tc", "simt"])
def test_conv_fwd_fused_finalize(geom, impl):
    """dp_conv_fwd_fin (BatchNorm finalisation by the conv kernel's last CTA) == dp_conv_fwd + dp_bn_finalize, bit for
    bit, in every statistics mode of the tcgen05 kernel and through the CUDA-core composition; the ticket word returns
    to zero, so the same word serves the next launch."""
    lib = L.load()
    Cc, K, k, s, p, T, H, W = geom
    big = H >= 64
    B = 2 if big else 3
    if big and impl == "simt":
        pytest.skip("CUDA-core family at full resolution is covered by the small shapes")
    dtype = torch.bfloat16
    x, w = make_case(geom, B, dtype, 3)
    xi = to_int(x, dtype)
    gm = Fn.conv_geom(Cc, K, k, s, p, xi)
    d = gm.desc
    wf = torch.empty((d.Kp, gm.taps, d.Cp), dtype=dtype, device=DEV)
    L.check(lib.dp_pack_weights(C.byref(d), w.contiguous().data_ptr(), wf.data_ptr(), None, L.stream_ptr()), "pack")
    im = L.IMPL_TC if impl == "tc" else L.IMPL_SIMT
    g = torch.Generator().manual_seed(4)
    gamma = (torch.rand(K, generator=g) + 0.5).to(DEV)
    beta = torch.randn(K, generator=g).to(DEV)
    # reference: partials + stand-alone finalize
    y0 = torch.empty(gm.out_shape, dtype=dtype, device=DEV)
    part = torch.zeros((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=DEV)
    nparts = C.c_int(0)
    L.check(lib.dp_conv_fwd(C.byref(d), xi.data_ptr(), wf.data_ptr(), y0.data_ptr(), part.data_ptr(), C.byref(nparts), im,
                            L.stream_ptr()), "fwd")
    st0 = torch.full((4, d.Kp), float("nan"), device=DEV)
    rm0, rv0 = torch.zeros(K, device=DEV), torch.ones(K, device=DEV)
    L.check(lib.dp_bn_finalize(part.data_ptr(), nparts.value, K, d.Kp, float(gm.rows_out), gamma.data_ptr(), beta.data_ptr(),
                               1e-5, 0.1, rm0.data_ptr(), rv0.data_ptr(), st0[0].data_ptr(), st0[1].data_ptr(),
                               st0[2].data_ptr(), st0[3].data_ptr(), L.stream_ptr()), "finalize")
    ticket = torch.zeros(1, dtype=torch.int32, device=DEV)
    for rep in range(2):   # twice on the same ticket word
        y1 = torch.empty_like(y0)
        st1 = torch.full((4, d.Kp), float("nan"), device=DEV)
        rm1, rv1 = torch.zeros(K, device=DEV), torch.ones(K, device=DEV)
        part1 = torch.zeros_like(part)
        fin = _fin_fwd(d, gm.rows_out, gamma, beta, rm1, rv1, st1, ticket)
        L.check(lib.dp_conv_fwd_fin(C.byref(d), xi.data_ptr(), wf.data_ptr(), y1.data_ptr(), part1.data_ptr(), C.byref(fin),
                                    im, L.stream_ptr()), "fwd_fin")
        torch.cuda.synchronize()
        assert int(ticket.item()) == 0
        assert torch.equal(y0, y1)
        assert torch.equal(st0, st1), (st0 - st1).abs().max()
        assert torch.equal(rm0, rm1) and torch.equal(rv0, rv1)
    # and the statistics themselves against fp64 on the bf16-rounded output
    yf = from_int(y0, K).double()
    mu = yf.mean(dim=(0, 2, 3, 4))
    assert float((st0[0, :K].double() - mu).abs().max() / yf.abs().max()) < 2e-3


@pytest.mark.parametrize("C_", [32, 72, 288])
@pytest.mark.parametrize("with_out", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_bn_bwd_reduce_fused_finalize(C_, with_out, training):
    """dp_bn_act_bwd_reduce_fin == dp_bn_act_bwd_reduce + dp_bn_bwd_finalize (bit for bit); coef_zero for eval-mode BN."""
    lib = L.load()
    Cp = Fn.ceil16(C_)
    rows = 3 * 5 * 16 * 16
    g = torch.Generator().manual_seed(8)
    mk = lambda: (torch.randn(rows, Cp, generator=g).to(DEV) * (torch.arange(Cp, device=DEV) < C_)).bfloat16()
    dz, y, out = mk(), mk(), mk()
    ss = torch.zeros((4, Cp), device=DEV)          # mean, rstd, scale, shift
    ss[0, :C_] = torch.randn(C_, generator=g).to(DEV) * 0.2
    ss[1, :C_] = torch.rand(C_, generator=g).to(DEV) + 0.5
    ss[2, :C_] = torch.rand(C_, generator=g).to(DEV) + 0.5
    ss[3, :C_] = torch.randn(C_, generator=g).to(DEV) * 0.3
    po = out.data_ptr() if with_out else None
    part = torch.zeros((L.DP_MAX_PARTS, 2, Cp), dtype=torch.float32, device=DEV)
    nparts = C.c_int(0)
    L.check(lib.dp_bn_act_bwd_reduce(dz.data_ptr(), y.data_ptr(), po, ss[2].data_ptr(), ss[3].data_ptr(), ss[0].data_ptr(),
                                     ss[1].data_ptr(), 0.01, 0.2, part.data_ptr(), C.byref(nparts), rows, Cp, L.DP_BF16,
                                     L.stream_ptr()), "reduce")
    dg0, db0, cf0 = torch.empty(C_, device=DEV), torch.empty(C_, device=DEV), torch.empty((2, Cp), device=DEV)
    L.check(lib.dp_bn_bwd_finalize(part.data_ptr(), nparts.value, C_, Cp, float(rows), ss[0].data_ptr(), ss[1].data_ptr(),
                                   dg0.data_ptr(), db0.data_ptr(), cf0.data_ptr(), L.stream_ptr()), "finalize")
    if not training:
        cf0.zero_()
    ticket = torch.zeros(1, dtype=torch.int32, device=DEV)
    for rep in range(2):
        dg1, db1 = torch.full((C_,), float("nan"), device=DEV), torch.full((C_,), float("nan"), device=DEV)
        cf1 = torch.full((2, Cp), float("nan"), device=DEV)
        part1 = torch.zeros_like(part)
        fin = L.BnFin(kind=2, C=C_, Cp=Cp, coef_zero=0 if training else 1, count=float(rows), gamma=None, beta=None, eps=0.0,
                      momentum=0.0, running_mean=None, running_var=None, mean=ss[0].data_ptr(), rstd=ss[1].data_ptr(),
                      scale=None, shift=None, dgamma=dg1.data_ptr(), dbeta=db1.data_ptr(), coef=cf1.data_ptr(),
                      ticket=ticket.data_ptr())
        L.check(lib.dp_bn_act_bwd_reduce_fin(dz.data_ptr(), y.data_ptr(), po, ss[2].data_ptr(), ss[3].data_ptr(), 0.01, 0.2,
                                             part1.data_ptr(), rows, Cp, L.DP_BF16, C.byref(fin), L.stream_ptr()), "reduce_fin")
        torch.cuda.synchronize()
        assert int(ticket.item()) == 0
        assert torch.equal(dg0, dg1) and torch.equal(db0, db1) and torch.equal(cf0, cf1)


@pytest.mark.parametrize("geom", [(72, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), 9, 16, 16),      # sums in the tcgen05 epilogue
                                  (32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), 5, 16, 16),
                                  (32, 115, (1, 3, 3), (1, 2, 2), (0, 1, 1), 4, 16, 16)],    # composed: dgrad + reduction
                         ids=lambda g: f"{g[0]}to{g[1]}_k{g[2]}_s{g[3]}")
def test_dgrad_bnstats_fused_finalize(geom):
    """dp_conv_dgrad_bnstats_fin == dp_conv_dgrad_bnstats + dp_bn_bwd_finalize of the producing layer (bit for bit)."""
    lib = L.load()
    B = 3
    x, w = make_case(geom, B, torch.bfloat16, 5)
    Cc, K, k, s, p, T, H, W = geom
    xi = to_int(x, torch.bfloat16)
    gm = Fn.conv_geom(Cc, K, k, s, p, xi)
    d = gm.desc
    wd = torch.empty((d.Cp, gm.taps, d.Kp), dtype=torch.bfloat16, device=DEV)
    L.check(lib.dp_pack_weights(C.byref(d), w.contiguous().data_ptr(), None, wd.data_ptr(), L.stream_ptr()), "pack")
    g = torch.Generator(device="cpu").manual_seed(12)
    dyi = torch.randn(gm.out_shape, generator=g).to(DEV).bfloat16()
    dyi[..., K:] = 0
    yprev = torch.randn(xi.shape, generator=g).to(DEV).bfloat16()
    yprev[..., Cc:] = 0
    ss = torch.zeros((4, d.Cp), dtype=torch.float32, device=DEV)   # producer's mean, rstd, scale, shift
    ss[0, :Cc] = torch.randn(Cc, generator=g).to(DEV) * 0.2
    ss[1, :Cc] = torch.rand(Cc, generator=g).to(DEV) + 0.5
    ss[2, :Cc] = torch.rand(Cc, generator=g).to(DEV) + 0.5
    ss[3, :Cc] = torch.randn(Cc, generator=g).to(DEV) * 0.3
    rows = gm.rows_in
    dx0 = torch.empty_like(xi)
    part = torch.zeros((L.DP_MAX_PARTS, 2, d.Cp), dtype=torch.float32, device=DEV)
    nparts = C.c_int(0)
    L.check(lib.dp_conv_dgrad_bnstats(C.byref(d), dyi.data_ptr(), wd.data_ptr(), None, dx0.data_ptr(), yprev.data_ptr(),
                                      ss[2].data_ptr(), 0.01, part.data_ptr(), C.byref(nparts), L.IMPL_AUTO, L.stream_ptr()),
            "dgrad_bnstats")
    dg0, db0, cf0 = torch.empty(Cc, device=DEV), torch.empty(Cc, device=DEV), torch.empty((2, d.Cp), device=DEV)
    L.check(lib.dp_bn_bwd_finalize(part.data_ptr(), nparts.value, Cc, d.Cp, float(rows), ss[0].data_ptr(), ss[1].data_ptr(),
                                   dg0.data_ptr(), db0.data_ptr(), cf0.data_ptr(), L.stream_ptr()), "finalize")
    ticket = torch.zeros(1, dtype=torch.int32, device=DEV)
    dg1, db1 = torch.full((Cc,), float("nan"), device=DEV), torch.full((Cc,), float("nan"), device=DEV)
    cf1 = torch.full((2, d.Cp), float("nan"), device=DEV)
    dx1 = torch.empty_like(xi)
    part1 = torch.zeros_like(part)
    fin = L.BnFin(kind=2, C=Cc, Cp=d.Cp, coef_zero=0, count=float(rows), gamma=None, beta=None, eps=0.0, momentum=0.0,
                  running_mean=None, running_var=None, mean=ss[0].data_ptr(), rstd=ss[1].data_ptr(), scale=None, shift=None,
                  dgamma=dg1.data_ptr(), dbeta=db1.data_ptr(), coef=cf1.data_ptr(), ticket=ticket.data_ptr())
    L.check(lib.dp_conv_dgrad_bnstats_fin(C.byref(d), dyi.data_ptr(), wd.data_ptr(), None, dx1.data_ptr(), yprev.data_ptr(),
                                          ss[2].data_ptr(), 0.01, part1.data_ptr(), C.byref(fin), L.IMPL_AUTO, L.stream_ptr()),
            "dgrad_bnstats_fin")
    torch.cuda.synchronize()
    assert int(ticket.item()) == 0
    assert torch.equal(dx0, dx1)
    assert torch.equal(dg0, dg1) and torch.equal(db0, db1) and torch.equal(cf0, cf1)
