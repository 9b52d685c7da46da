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
def cell(r, conv, col):
    if r is None:
        return "—"
    t = r[1]
    hb = r[4] * 1e6 / (HBM * 1e9) * 1e6
    bound = max(r[2] * t / TF, hb) if conv else hb
    a = tot.setdefault(col, [0.0, 0.0]); a[0] += t; a[1] += bound
    return f"{t:.0f} µs ({bound / t * 100:.0f} %)"
cols = ["fwd", "apply", "dgrad", "wgrad", "reduce", "bwd_apply"]
lines = ["| layer | fwd conv | BN apply | dgrad | wgrad | BN-bwd reduce | BN-bwd apply |", "|---|---|---|---|---|---|---|"]
for n in fwd:
    s = stats[n]
    cells = [cell(s.get(c), c in ("fwd", "dgrad", "wgrad"), c) for c in cols]
    if s.get("reduce") is None:
        cells[4] = "in the dgrad epilogue"
    if s.get("dgrad") is None:
        cells[2] = "— (no data gradient)"
    lines.append(f"| {n.replace('res2plus1d.', '')} | " + " | ".join(cells) + " |")
lines.append("| **sum** | " + " | ".join(f"**{tot[c][0] / 1e3:.2f} ms ({tot[c][1] / tot[c][0] * 100:.0f} %)**" for c in cols) + " |")
open(dst, "w").write(f"""# Per-layer roofline of one training step (B = 64, one B200, per-kernel CUDA events, eager launches)

Source: `DP_BENCH_DUMP={src} python bench.py`, table by `scripts/layer_roofline.py`.  Each cell: kernel time and, in
parentheses, the fraction of its roofline bound it achieves — conv kernels `max(algorithmic FLOPs / {TF} TFLOP/s,
algorithmic bytes / {HBM} GB/s) / time`, BN passes `algorithmic bytes / {HBM} GB/s / time` (peaks:
`MEASURED_PEAKS.json`, sustained bf16 and copy bandwidth).  Algorithmic bytes = every operand tensor once.  A dgrad that
also produces the BatchNorm-backward sums of the layer above it reads that layer's conv output as well (not counted).

""" + "\n".join(lines) + "\n")
print("\n".join(lines[-6:]))
