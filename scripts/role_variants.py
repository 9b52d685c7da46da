"""Kernel-level A/B runs of the tcgen05 gather kernel on the stage-1 layer shapes (development aid)."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dp_b200
from dp_b200 import _lib as L, functional as Fn
B = 64
lib = L.load(); L.require_device()
dev = "cuda"
LAYERS = [
    ("stem.temporal 45->32", 45, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), (21, 64, 64)),
    ("conv2.spatial 32->72", 32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), (21, 64, 64)),
    ("conv2.temporal 72->32", 72, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), (21, 64, 64)),
    ("conv3.spatial 64->144", 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), (11, 32, 32)),
    ("conv3.temporal 144->64", 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), (11, 32, 32)),
]
VARIANTS = [eval(v) for v in os.environ.get("VARIANTS", "{}").split(";")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(name, fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"  {name:40s} {min(ts):8.1f} us", flush=True)
for (name, cin, cout, k, s, p, inp) in LAYERS:
    print(name, flush=True)
    x = torch.randn(B, *inp, Fn.ceil16(cin), device=dev).bfloat16(); x[..., cin:] = 0
    gm = Fn.conv_geom(cin, cout, k, s, p, x); d = gm.desc
    w = torch.randn(cout, cin, *k, device=dev)
    wf, wd = Fn.pack_weights(w, gm, torch.bfloat16, None)
    y = torch.empty(gm.out_shape, dtype=torch.bfloat16, device=dev)
    part = torch.empty((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=dev); nparts = C.c_int(0)
    st = L.stream_ptr()
    dy = torch.randn(gm.out_shape, device=dev).bfloat16(); dy[..., cout:] = 0
    dx = torch.empty_like(x)
    for opts in VARIANTS:
        for kk, v in opts.items(): L.set_option(kk, v[0])
        run("fwd   " + str(opts), lambda: L.check(lib.dp_conv_fwd(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(), C.byref(nparts), 0, st)))
        run("fwd nostats " + str(opts), lambda: L.check(lib.dp_conv_fwd(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), None, None, 0, st)))
        run("dgrad " + str(opts), lambda: L.check(lib.dp_conv_dgrad(C.byref(d), dy.data_ptr(), wd.data_ptr(), None, dx.data_ptr(), 0, st)))
        for kk, v in opts.items(): L.set_option(kk, v[1])
