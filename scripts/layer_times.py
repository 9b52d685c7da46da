"""Per-layer timing of the conv kernels (fwd / dgrad / wgrad) and BN kernels at the bench shape, CUDA events."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dp_b200
from dp_b200 import _lib as L, functional as Fn
from oracle import r2plus1d_port as port

B = int(os.environ.get("B", "64"))
REP = int(os.environ.get("REP", "5"))
lib = L.load(); L.require_device()
dev = "cuda"
layers = port.all_conv_layers([1, 2, 2, 1], 1.0)
# spatial sizes: replay the network
def out_dim(n, k, s, p): return (n + 2 * p - k) // s + 1
shapes = {}
T, H, W = 21, 128, 128
stem, blocks = port.encoder_plan([1, 2, 2, 1], 1.0)
cur = (T, H, W)
seq = []
def run_seq(ls, inp):
    for (name, cin, cout, k, s, p, slope) in ls:
        seq.append((name, cin, cout, k, s, p, inp))
        inp = tuple(out_dim(inp[i], k[i], s[i], p[i]) for i in range(3))
    return inp
cur = run_seq(stem, cur)
for b in blocks:
    o = run_seq(b["conv1"], cur); o = run_seq(b["conv2"], o)
    if b["shortcut"]: run_seq(b["shortcut"], cur)
    cur = o
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(REP):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
tot = {"fwd": 0, "dgrad": 0, "wgrad": 0}
print(f"{'layer':52s} {'in(T,H,W)':>12s} {'fwd ms':>8s} {'TF/s':>6s} {'GB/s':>6s} | {'dgrad':>7s} {'TF/s':>6s} | {'wgrad':>7s} {'TF/s':>6s} {'GB/s':>6s}")
for (name, cin, cout, k, s, p, inp) in seq:
    x = torch.randn(B, *inp, Fn.ceil16(cin), device=dev).bfloat16()
    x[..., cin:] = 0
    gm = Fn.conv_geom(cin, cout, k, s, p, x)
    d = gm.desc
    w = torch.randn(cout, cin, *k, device=dev)
    wf, wd = Fn.pack_weights(w, gm, torch.bfloat16, None)
    y = torch.empty(gm.out_shape, dtype=torch.bfloat16, device=dev)
    part = torch.empty((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=dev)
    nparts = C.c_int(0)
    st = L.stream_ptr()
    def f_fwd(): L.check(lib.dp_conv_fwd(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(), C.byref(nparts), 0, st))
    dy = torch.randn(gm.out_shape, device=dev).bfloat16(); dy[..., cout:] = 0
    dx = torch.empty_like(x)
    def f_dg(): L.check(lib.dp_conv_dgrad(C.byref(d), dy.data_ptr(), wd.data_ptr(), None, dx.data_ptr(), 0, st))
    dw = torch.empty_like(w)
    ws = torch.empty(max(16, int(lib.dp_conv_wgrad_workspace(C.byref(d), 0))), dtype=torch.uint8, device=dev)
    def f_wg(): L.check(lib.dp_conv_wgrad(C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(), 0, st))
    t_f, t_d, t_w = timeit(f_fwd), timeit(f_dg), timeit(f_wg)
    fl = gm.flops; by = 2.0 * (gm.rows_in * cin + gm.rows_out * cout)
    tot["fwd"] += t_f; tot["dgrad"] += t_d; tot["wgrad"] += t_w
    print(f"{name[11:]:52s} {str(inp):>12s} {t_f:8.3f} {fl/t_f/1e9:6.0f} {by/t_f/1e6:6.0f} | {t_d:7.3f} {fl/t_d/1e9:6.0f} | {t_w:7.3f} {fl/t_w/1e9:6.0f} {by/t_w/1e6:6.0f}", flush=True)
    del x, y, dy, dx, ws
print("totals ms", tot)
