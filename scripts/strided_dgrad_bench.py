"""Strided data gradients of the R(2+1)D down-sampling blocks at B = 64: one launch per stride-parity class
(dp_conv_dgrad) against every class in one launch (dp_conv_dgrad_classes).  Development aid / evidence for profiles/."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DP_DEBUG_PLAN", "0")
import dp_b200
from dp_b200 import _lib as L, functional as Fn
B = int(os.environ.get("B", "64"))
lib = L.load(); L.require_device()
dev = "cuda"
LAYERS = [
    ("conv3.b1.c1.spatial 32->115 s(1,2,2)", 32, 115, (1, 3, 3), (1, 2, 2), (0, 1, 1), (21, 64, 64)),
    ("conv3.b1.c1.temporal 115->64 s(2,1,1)", 115, 64, (3, 1, 1), (2, 1, 1), (1, 0, 0), (21, 32, 32)),
    ("conv3.b1.ds.spatial 32->21 1x1 s(1,2,2)", 32, 21, (1, 1, 1), (1, 2, 2), (0, 0, 0), (21, 64, 64)),
    ("conv3.b1.ds.temporal 21->64 1x1 s(2,1,1)", 21, 64, (1, 1, 1), (2, 1, 1), (0, 0, 0), (21, 32, 32)),
    ("conv4.b1.c1.spatial 64->144 s(1,2,2)", 64, 144, (1, 3, 3), (1, 2, 2), (0, 1, 1), (11, 32, 32)),
    ("conv4.b1.c1.temporal 144->64 s(2,1,1)", 144, 64, (3, 1, 1), (2, 1, 1), (1, 0, 0), (11, 16, 16)),
    ("conv4.b1.ds.spatial 64->32 1x1 s(1,2,2)", 64, 32, (1, 1, 1), (1, 2, 2), (0, 0, 0), (11, 32, 32)),
    ("conv5.b1.c1.spatial 64->230 s(1,2,2)", 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1), (6, 16, 16)),
    ("conv5.b1.c1.temporal 230->128 s(2,1,1)", 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0), (6, 8, 8)),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)
tot = [0.0, 0.0]
for (name, cin, cout, k, s, p, inp) in LAYERS:
    x = torch.zeros(B, *inp, Fn.ceil16(cin), device=dev, dtype=torch.bfloat16)
    gm = Fn.conv_geom(cin, cout, k, s, p, x); d = gm.desc
    w = torch.randn(cout, cin, *k, device=dev)
    wd = torch.empty((d.Cp, gm.taps, d.Kp), dtype=torch.bfloat16, device=dev)
    L.check(lib.dp_pack_weights(C.byref(d), w.data_ptr(), None, wd.data_ptr(), L.stream_ptr()))
    dy = torch.randn(gm.out_shape, device=dev).bfloat16(); dy[..., cout:] = 0
    ad = torch.randn_like(x) if (k == (1, 3, 3)) else None     # only a block's first conv receives the shortcut gradient
    dx = torch.empty_like(x)
    st = L.stream_ptr()
    t_a = run(lambda: L.check(lib.dp_conv_dgrad(C.byref(d), dy.data_ptr(), wd.data_ptr(), (ad.data_ptr() if ad is not None else None), dx.data_ptr(), L.IMPL_TC, st)))
    ref = dx.clone()
    n = int(lib.dp_dgrad_classes_weight_elems(C.byref(d), L.IMPL_TC))
    if n:
        wc = torch.empty(n, dtype=torch.bfloat16, device=dev)
        L.check(lib.dp_pack_weights_dgrad_classes(C.byref(d), w.data_ptr(), wc.data_ptr(), st))
        os.environ["DP_DEBUG_PLAN"] = "1"
        t_b = run(lambda: L.check(lib.dp_conv_dgrad_classes(C.byref(d), dy.data_ptr(), wc.data_ptr(), (ad.data_ptr() if ad is not None else None), dx.data_ptr(), st)))
        same = bool(torch.equal(ref, dx)) or float((ref.float() - dx.float()).abs().max())
    else:
        t_b, same = float("nan"), "n/a"
    mb = (dy.numel() + 2 * x.numel()) * 2 / 1e6
    print(f"{name:44s} per-class launches {t_a:8.1f} us | one launch {t_b:8.1f} us | algorithmic {mb:7.1f} MB -> {mb / max(t_b, 1e-9) :6.2f} TB/s... equal={same}", flush=True)
    tot[0] += t_a; tot[1] += t_b if n else t_a
print(f"sum: {tot[0]:.1f} us -> {tot[1]:.1f} us")
