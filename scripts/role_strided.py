"""Per-role cycle accounting of the one-launch strided data gradient (development aid)."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dp_b200
from dp_b200 import _lib as L, functional as Fn
B = 64
lib = L.load(); L.require_device()
dev = "cuda"
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("="); L.set_option(k, int(v))
LAYERS = [
    ("conv3.b1.c1.spatial 32->115 s(1,2,2)", 32, 115, (1, 3, 3), (1, 2, 2), (0, 1, 1), (21, 64, 64)),
    ("conv3.b1.c1.temporal 115->64 s(2,1,1)", 115, 64, (3, 1, 1), (2, 1, 1), (1, 0, 0), (21, 32, 32)),
]
def run(name, fn):
    fn(); torch.cuda.synchronize()
    dbg.zero_()
    lib.dp_set_debug_buffer(dbg.data_ptr(), dbg.numel() * 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.dp_set_debug_buffer(None, 0)
    d2 = dbg[148 * 8:].view(148, 8).double().mean(0)
    d = dbg[:148 * 8].view(148, 8).double()
    act = d[:, 1] > 0
    m = d[act].mean(0)
    print(f"  {name:22s} {e0.elapsed_time(e1)*1e3:7.1f}us |prod wait {m[0]/1e3:5.0f}k/{m[1]/1e3:5.0f}k |mma wfull {m[2]/1e3:5.0f}k wtmem {m[3]/1e3:5.0f}k /{m[4]/1e3:5.0f}k |epi wtfull {m[5]/1e3:5.0f}k /{m[6]/1e3:5.0f}k | epi: waitfree {d2[0]/1e3:.0f}k drain {d2[1]/1e3:.0f}k fence+bar {d2[2]/1e3:.0f}k rest {d2[3]/1e3:.0f}k", flush=True)
for (name, cin, cout, k, s, p, inp) in LAYERS:
    print(name, flush=True)
    x = torch.zeros(B, *inp, Fn.ceil16(cin), device=dev, dtype=torch.bfloat16)
    gm = Fn.conv_geom(cin, cout, k, s, p, x); d = gm.desc
    w = torch.randn(cout, cin, *k, device=dev)
    dy = torch.randn(gm.out_shape, device=dev).bfloat16(); dy[..., cout:] = 0
    ad = torch.randn_like(x); dx = torch.empty_like(x)
    st = L.stream_ptr()
    n = int(lib.dp_dgrad_classes_weight_elems(C.byref(d), L.IMPL_TC))
    wc = torch.empty(n, dtype=torch.bfloat16, device=dev)
    L.check(lib.dp_pack_weights_dgrad_classes(C.byref(d), w.data_ptr(), wc.data_ptr(), st))
    for skip in (0, 1, 2, 4):
        L.set_option("tc_dbg_skip", skip)
        run(f"skip={skip} no addend", lambda: L.check(lib.dp_conv_dgrad_classes(C.byref(d), dy.data_ptr(), wc.data_ptr(), None, dx.data_ptr(), st)))
        if skip == 0:
            run(f"skip={skip} addend", lambda: L.check(lib.dp_conv_dgrad_classes(C.byref(d), dy.data_ptr(), wc.data_ptr(), ad.data_ptr(), dx.data_ptr(), st)))
    L.set_option("tc_dbg_skip", 0)
