"""Forward conv (with BatchNorm statistics) of the wide spatial layers at the bench batch under the two values of one
load-time option (OPT=tc_nsplit: one launch with weights streamed per tile against the output-channel split with
resident weights; OPT=tc_stats_keep: staged-tile statistics combined per tile against once per CTA).  CUDA events, L2
flushed between launches."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dp_b200
from dp_b200 import _lib as L, functional as Fn

B = int(os.environ.get("B", "64"))
REP = int(os.environ.get("REP", "9"))
lib = L.load(); L.require_device()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(REP):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


OPT = os.environ.get("OPT", "tc_nsplit")     # the option whose values 0 / 1 are compared


def fwd_time(cin, cout, k, s, p, inp, nsplit):
    L.set_option(OPT, nsplit)
    x = torch.randn(B, *inp, Fn.ceil16(cin), device=dev).bfloat16()
    gm = Fn.conv_geom(cin, cout, k, s, p, x)
    d = gm.desc
    w = torch.randn(cout, cin, *k, device=dev)
    wf, wd = Fn.pack_weights(w, gm, torch.bfloat16, None)
    y = torch.empty(gm.out_shape, dtype=torch.bfloat16, device=dev)
    part = torch.empty((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=dev)
    nparts = C.c_int(0)
    st = L.stream_ptr()
    buf = C.create_string_buffer(1024)
    lib.dp_conv_describe_plan(C.byref(d), 0, 1, buf, 1024)
    t = timeit(lambda: L.check(lib.dp_conv_fwd(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(), C.byref(nparts), 0, st)))
    plan = buf.value.decode()
    keys = ("MT=", "Ntile=", "n_ntiles=", "resident=", "reg_stats=", "mma_stats=", "dual=", "stages=", "lps=", "split=")
    return t, " ".join(tok for tok in plan.split() if tok.startswith(keys))


for name, cin, cout, k, s, p, inp in (
        ("conv3 spatial 64->144", 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), (11, 32, 32)),
        ("   first half 64->80 ", 64, 80, (1, 3, 3), (1, 1, 1), (0, 1, 1), (11, 32, 32)),
        ("  second half 64->64 ", 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), (11, 32, 32)),
        ("conv4 spatial 64->144", 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), (6, 16, 16)),
        ("conv4 strided 64->144", 64, 144, (1, 3, 3), (1, 2, 2), (0, 1, 1), (11, 32, 32)),
        ("conv5 spatial 128->288", 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), (3, 8, 8))):
    for ns in (0, 1):
        t, plan = fwd_time(cin, cout, k, s, p, inp, ns)
        print(f"{name:24s} {OPT}={ns}  {t:8.1f} us   {plan}", flush=True)
