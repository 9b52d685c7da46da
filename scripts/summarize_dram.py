"""gpurun_out/<tag>_step_launches.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over
the NVTX range of ONE eager training step) -> profiles/<tag>_step_launches.txt (every launch of the step, in order),
profiles/<tag>_launches_summary.txt (per-kernel shares) and profiles/<tag>_dram_per_launch.json (mean DRAM bytes per
launch per kernel family over ALL its launches of the step: what bench.py prints as roofline.traffic).
If gpurun_out/<tag>_prof_step_raw.csv exists (ncu --set full of the same step) the key metrics of every captured launch
go to profiles/<tag>_prof_step_raw.csv."""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
src, out = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rows = [r for r in csv.reader(open(os.path.join(src, f"{tag}_step_launches.csv"))) if len(r) > 5]
hdr = rows[0]
ki, mi, ui, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
launch = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    except ValueError:
        continue
    d = launch.setdefault(r[idi], {"name": r[ki]})
    d[r[mi]] = v


def family(name):
    if "tc_gather_gemm" in name: return "tc_gather_gemm"
    if "wgrad_tc_kernel" in name: return "tc_wgrad"
    if "bn_act_bwd_apply" in name: return "bn_act_bwd_apply"
    if "bn_act_apply" in name: return "bn_act_apply"
    if "col_reduce" in name or "bwd_reduce" in name: return "bn_act_bwd_reduce"
    return name.split("(")[0].split("<")[0].replace("void ", "").replace("dp::", "")[-48:]


fam = collections.defaultdict(lambda: {"launches": 0, "us": 0.0, "dram": 0.0})
with open(os.path.join(out, f"{tag}_step_launches.txt"), "w") as f:
    f.write("every kernel of ONE eager training step (B = 64, bf16), in launch order: ncu --profile-from-start off (bench.py --nvtx-step brackets one step with cudaProfilerStart/Stop) "
            "--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n"
            "(durations under ncu are cold-cache and serialised: compare SHARES, not absolutes)\n\n")
    f.write(f"{'#':>4s} {'us':>9s} {'read MB':>9s} {'write MB':>9s}  kernel\n")
    for i, (k, d) in enumerate(launch.items()):
        us, rd, wr = d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        f.write(f"{i:4d} {us:9.1f} {rd/1e6:9.1f} {wr/1e6:9.1f}  {d['name'].split('(')[0][-90:]}\n")
        a = fam[family(d["name"])]
        a["launches"] += 1; a["us"] += us; a["dram"] += rd + wr
tot = sum(a["us"] for a in fam.values())
with open(os.path.join(out, f"{tag}_launches_summary.txt"), "w") as f:
    f.write(f"{len(launch)} launches of one eager step, {tot/1e3:.3f} ms total under ncu (cold-cache, serialised: compare SHARES)\n\n")
    f.write(f"{'ms':>9s} {'share':>6s} {'n':>5s} {'DRAM GB':>8s}  family\n")
    for n, a in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        f.write(f"{a['us']/1e3:9.3f} {100*a['us']/tot:5.1f}% {a['launches']:5d} {a['dram']/1e9:8.3f}  {n}\n")
json.dump({"source": f"{tag}_step_launches.csv", "families": {n: {"launches": a["launches"], "dram_bytes_per_launch": a["dram"] / a["launches"],
                                                                  "us_per_launch_under_ncu": a["us"] / a["launches"]}
                                                              for n, a in fam.items()}},
          open(os.path.join(out, f"{tag}_dram_per_launch.json"), "w"), indent=1)
raw = os.path.join(src, f"{tag}_prof_step_raw.csv")
if not os.path.exists(raw) and os.path.exists(raw + ".gz"):
    import gzip, shutil
    with gzip.open(raw + ".gz", "rb") as fi, open(raw, "wb") as fo:
        shutil.copyfileobj(fi, fo)
if os.path.exists(raw):
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_tc.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
    rr = [r for r in csv.reader(open(raw)) if len(r) > 10]
    h = rr[0]
    idx = [i for i, c in enumerate(h) if c in want]
    with open(os.path.join(out, f"{tag}_prof_step_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in idx])
        for r in rr[1:]:
            w.writerow([r[i][:70] for i in idx])
print("wrote", sorted(x for x in os.listdir(out) if x.startswith(tag)))
