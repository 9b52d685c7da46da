"""Turn a DP_BENCH_DUMP per-launch list of one training step into the per-layer roofline table under profiles/.
usage: python scripts/layer_roofline.py gpurun_out/step_dump_final.txt profiles/r1e_layer_roofline.md"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r2plus1d_port as port

src, dst = sys.argv[1], sys.argv[2]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM, TF = peaks["hbm_gbs"], peaks["bf16_tflops_sustained"]
stem, blocks = port.encoder_plan([1, 2, 2, 1], 1.0)
fwd = [l[0] for l in stem]
for b in blocks:
    c1 = [l[0] for l in b["conv1"]]; c2 = [l[0] for l in b["conv2"]]
    sc = [l[0] for l in b["shortcut"]] if b["shortcut"] else []
    fwd += c1 + [c2[0]] + sc + [c2[1]]          # call order inside ResBlockFn.forward
# ---- tensor-pipe OPERAND bound (profiles/ubench/mma_rate.log): a tcgen05.mma with both operands in shared memory costs
# max(math floor, operand read at ~128 B/clk) cycles -- M128 x N x K16: N=32 43.6, 64 53.1, 80 57.8, 128 72.1, 256 136 --
# so a layer with few channels is bound by the A-operand (activation tile) reads, not by FLOPs or HBM.
UB_N = [16, 32, 64, 80, 128, 256]
UB_C = [41.3, 43.6, 53.1, 57.8, 72.1, 136.1]
def mma_cycles(n):
    n = max(16, min(256, n))
    for i in range(len(UB_N) - 1):
        if n <= UB_N[i + 1]:
            f = (n - UB_N[i]) / (UB_N[i + 1] - UB_N[i])
            return UB_C[i] + f * (UB_C[i + 1] - UB_C[i])
    return UB_C[-1]
SMS, CLK = 148, peaks.get("sm_max_mhz", 1965.0) * 1e6
c16 = lambda c: (c + 15) // 16 * 16
def ntile(n):
    nt = (n + 255) // 256
    return nt, n // nt
B = 64
geo = {}    # name -> (cin, cout, taps, stride product, rows_in, rows_out)
def walk(layers, dims):
    for (name, cin, cout, k, s, p, _) in layers:
        T, H, W = dims
        o = ((T + 2 * p[0] - k[0]) // s[0] + 1, (H + 2 * p[1] - k[1]) // s[1] + 1, (W + 2 * p[2] - k[2]) // s[2] + 1)
        geo[name] = (cin, cout, k[0] * k[1] * k[2], s[0] * s[1] * s[2], B * T * H * W, B * o[0] * o[1] * o[2])
        dims = o
    return dims
dims = walk(stem, (21, 128, 128))
for b in blocks:
    d1 = walk(b["conv1"], dims)
    if b["shortcut"]:
        walk(b["shortcut"], dims)
    dims = walk(b["conv2"], d1)
def operand_us(name, op):
    cin, cout, taps, sp, rin, rout = geo[name]
    Cp, Kp = c16(cin), c16(cout)
    if cin == 3:            # packed stem row pairs: 4 taps of K = 64
        taps, Cp = 4, 64
    if op == "fwd":
        nt, N = ntile(Kp)
        cyc = rout / 128 * taps * (Cp / 16) * nt * mma_cycles(N)
    elif op == "dgrad":     # every input pixel sees taps / stride-product taps
        nt, N = ntile(Cp)
        cyc = rin / 128 * (taps / sp) * (Kp / 16) * nt * mma_cycles(N)
    else:                   # wgrad: reduction over pixels in steps of 16, x windows stacked on M (128 rows), N = Kp
        groups = -(-taps * Cp // 128)
        nt, N = ntile(Kp)
        cyc = rout / 16 * groups * nt * mma_cycles(N)
    return cyc / SMS / CLK * 1e6
recs = []
for l in open(src):
    r = l.split()
    recs.append((r[0], float(r[1]), float(r[3]), float(r[5]), float(r[7])))   # family, us, TF/s, GB/s, MB
stats, i = {}, 0
for n in fwd:
    g, a = recs[i], recs[i + 1]; i += 2
    assert g[0] == "tc_gather_gemm" and a[0] == "bn_act_apply", (n, g, a)
    stats[n] = {"fwd": g, "apply": a}
groups, prev = [], None
for r in recs[i:]:
    if r[0] == "bn_act_bwd_reduce" or (r[0] == "bn_act_bwd_apply" and prev != "bn_act_bwd_reduce"):
        groups.append([])
    groups[-1].append(r); prev = r[0]
border = []
for b in reversed(blocks):                       # order of ResBlockFn.backward
    c1 = [l[0] for l in b["conv1"]]; c2 = [l[0] for l in b["conv2"]]
    sc = [l[0] for l in b["shortcut"]] if b["shortcut"] else []
    border += [c2[1], c2[0], c1[1]] + list(reversed(sc)) + [c1[0]]
border += [stem[1][0], stem[0][0]]
assert len(groups) == len(border), (len(groups), len(border))
key = {"bn_act_bwd_reduce": "reduce", "bn_act_bwd_apply": "bwd_apply", "tc_wgrad": "wgrad", "tc_gather_gemm": "dgrad"}
for n, gp in zip(border, groups):
    for r in gp:
        stats[n][key[r[0]]] = r
tot = {}
def cell(r, conv, col, name=None):
    if r is None:
        return "—"
    t = r[1]
    hb = r[4] * 1e6 / (HBM * 1e9) * 1e6
    bound = max(r[2] * t / TF, hb) if conv else hb
    a = tot.setdefault(col, [0.0, 0.0, 0.0]); a[0] += t; a[1] += bound
    if not conv:
        a[2] += bound
        return f"{t:.0f} µs ({bound / t * 100:.0f} %)"
    ob = max(bound, operand_us(name, col))
    a[2] += ob
    return f"{t:.0f} µs ({bound / t * 100:.0f} % / {ob / t * 100:.0f} %)"
cols = ["fwd", "apply", "dgrad", "wgrad", "reduce", "bwd_apply"]
lines = ["| layer | fwd conv | BN apply | dgrad | wgrad | BN-bwd reduce | BN-bwd apply |", "|---|---|---|---|---|---|---|"]
for n in fwd:
    s = stats[n]
    cells = [cell(s.get(c), c in ("fwd", "dgrad", "wgrad"), c, n) for c in cols]
    if s.get("reduce") is None:
        cells[4] = "in the dgrad epilogue"
    if s.get("dgrad") is None:
        cells[2] = "— (no data gradient)"
    lines.append(f"| {n.replace('res2plus1d.', '')} | " + " | ".join(cells) + " |")
lines.append("| **sum** | " + " | ".join(
    f"**{tot[c][0] / 1e3:.2f} ms ({tot[c][1] / tot[c][0] * 100:.0f} %" + (f" / {tot[c][2] / tot[c][0] * 100:.0f} %" if c in ("fwd", "dgrad", "wgrad") else "") + ")**"
    for c in cols) + " |")
step_t = sum(tot[c][0] for c in cols); step_b = sum(tot[c][2] for c in cols)
lines.append("")
lines.append(f"All six columns: {step_t / 1e3:.2f} ms of kernels against {step_b / 1e3:.2f} ms of summed per-layer bounds = {step_b / step_t * 100:.0f} %.")
open(dst, "w").write(f"""# Per-layer roofline of one training step (B = 64, one B200, per-kernel CUDA events, eager launches)

Source: `DP_BENCH_DUMP={src} python bench.py`, table by `scripts/layer_roofline.py`.  Each cell: kernel time and, in
parentheses, the fraction of its roofline bound it achieves — conv kernels `max(algorithmic FLOPs / {TF} TFLOP/s,
algorithmic bytes / {HBM} GB/s) / time`, BN passes `algorithmic bytes / {HBM} GB/s / time` (peaks:
`MEASURED_PEAKS.json`, sustained bf16 and copy bandwidth).  Algorithmic bytes = every operand tensor once.  A dgrad that
also produces the BatchNorm-backward sums of the layer above it reads that layer's conv output as well (not counted).
Conv cells carry a SECOND fraction: against `max(FLOPs, HBM, tensor-pipe operand rate)`, where the operand bound counts the
layer's `tcgen05.mma` instructions (128-pixel tiles x taps x channel steps of 16) at the measured cycles per instruction
with both operands in shared memory (`profiles/ubench/mma_rate.log`: M128 x N32 43.6, N64 53.1, N80 57.8, N128 72.1, N256 136
cycles -- the 4 KB activation operand is re-read for every instruction at ~128 B/clk, so 32-80 channel layers cannot reach
the FLOP or HBM roofline on this tensor pipe whatever the kernel does).

""" + "\n".join(lines) + "\n")
print("\n".join(lines[-6:]))
