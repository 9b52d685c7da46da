"""Per-role cycle accounting of the tcgen05 kernels on selected layers (development aid)."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DP_DEBUG_PLAN"] = "1"
import dp_b200
from dp_b200 import _lib as L, functional as Fn

B = int(os.environ.get("B", "64"))
lib = L.load(); L.require_device()
dev = "cuda"
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("="); L.set_option(k, int(v))
LAYERS = [
    ("stem.temporal 45->32", 45, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), (21, 64, 64)),
    ("conv2.spatial 32->72", 32, 72, (1, 3, 3), (1, 1, 1), (0, 1, 1), (21, 64, 64)),
    ("conv2.temporal 72->32", 72, 32, (3, 1, 1), (1, 1, 1), (1, 0, 0), (21, 64, 64)),
    ("conv3.spatial 64->144", 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), (11, 32, 32)),
    ("conv3.temporal 144->64", 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), (11, 32, 32)),
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
    print(f"  {name:6s} {e0.elapsed_time(e1)*1e3:7.1f}us |prod wait {m[0]/1e3:5.0f}k/{m[1]/1e3:5.0f}k |mma wfull {m[2]/1e3:5.0f}k wtmem {m[3]/1e3:5.0f}k /{m[4]/1e3:5.0f}k |epi wtfull {m[5]/1e3:5.0f}k /{m[6]/1e3:5.0f}k ({int(act.sum())} CTAs) | epi parts: waitfree+bar {d2[0]:.0f} drain {d2[1]:.0f} fence+bar {d2[2]:.0f} store {d2[3]:.0f}", flush=True)
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
    dx = torch.empty_like(x); dw = torch.empty_like(w)
    ws = torch.empty(max(16, int(lib.dp_conv_wgrad_workspace(C.byref(d), 0))), dtype=torch.uint8, device=dev)
    run("fwd", lambda: L.check(lib.dp_conv_fwd(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(), C.byref(nparts), 0, st)))
    run("dgrad", lambda: L.check(lib.dp_conv_dgrad(C.byref(d), dy.data_ptr(), wd.data_ptr(), None, dx.data_ptr(), 0, st)))
    run("wgrad", lambda: L.check(lib.dp_conv_wgrad(C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(), 0, st)))
