"""Per-role cycle accounting of the stem fast path (packed-rows 7x7/2 conv) forward and weight gradient."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DP_DEBUG_PLAN"] = "1"
import dp_b200
from dp_b200 import _lib as L, functional as Fn
B = int(os.environ.get("B", "64"))
lib = L.load(); L.require_device()
dev = "cuda"
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("="); L.set_option(k, int(v))
cfg = Fn.LayerCfg(3, 45, (1, 7, 7), (1, 2, 2), (0, 3, 3), 1.0)
geom = Fn.stem_geom(cfg, B, 21, 128, 128)
d = geom.desc
x = torch.randn(B, 3, 21, 128, 128, device=dev)
xp = Fn.stem_pack_input(x, geom)
w = torch.randn(45, 3, 1, 7, 7, device=dev)
wf, _ = Fn.pack_weights(w, geom, torch.bfloat16, None)
y = torch.empty(geom.out_shape, dtype=torch.bfloat16, device=dev)
part = torch.empty((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=dev); nparts = C.c_int(0)
dy = torch.randn(geom.out_shape, device=dev).bfloat16()
dw = torch.empty_like(w)
ws = torch.empty(max(16, int(lib.dp_stem_wgrad_workspace(C.byref(d)))), dtype=torch.uint8, device=dev)
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
st = L.stream_ptr()
def run(name, fn):
    fn(); torch.cuda.synchronize()
    dbg.zero_()
    lib.dp_set_debug_buffer(dbg.data_ptr(), dbg.numel() * 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.dp_set_debug_buffer(None, 0)
    d2 = dbg[148 * 8:].view(148, 8).double().mean(0)
    dd = dbg[:148 * 8].view(148, 8).double()
    m = dd[dd[:, 1] > 0].mean(0) if (dd[:, 1] > 0).any() else dd.mean(0)
    print(f"  {name:6s} {e0.elapsed_time(e1)*1e3:7.1f}us |prod wait {m[0]/1e3:5.0f}k/{m[1]/1e3:5.0f}k |mma wfull {m[2]/1e3:5.0f}k wtmem {m[3]/1e3:5.0f}k /{m[4]/1e3:5.0f}k |epi wtfull {m[5]/1e3:5.0f}k /{m[6]/1e3:5.0f}k | epi parts: waitfree {d2[0]:.0f} drain {d2[1]:.0f} fence+bar {d2[2]:.0f} store {d2[3]:.0f}", flush=True)
run("fwd", lambda: L.check(lib.dp_stem_conv_fwd(C.byref(d), xp.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(), C.byref(nparts), st)))
run("wgrad", lambda: L.check(lib.dp_stem_conv_wgrad(C.byref(d), xp.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(), st)))
