"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep} into the tracked summaries under profiles/."""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = os.path.join(ROOT, "profiles")
src = os.path.join(ROOT, "gpurun_out")
pre = (tag + "_") if os.path.exists(os.path.join(src, tag + "_launches.csv")) else ""
rows = [r for r in csv.reader(open(os.path.join(src, pre + "launches.csv"))) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    n = r[ki].split("(")[0][-70:]
    agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out, f"{tag}_launches_summary.txt"), "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 700 (python bench.py --steps 2 --warmup 3 --no-e2e --profile-steps 0 --no-cpu-baseline --no-graph)\n")
    f.write(f"{sum(v[0] for v in agg.values())} launches, {tot/1e6:.3f} ms total (cold-cache, serialised: compare SHARES)\n\n")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t/1e6:9.3f} ms {100*t/tot:5.1f}% {c:5d}  {n}\n")
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tc.sum"]
for name in ("prof_gather", "prof_wgrad", "prof_bn"):
    rep = os.path.join(src, pre + name + ".ncu-rep")
    if not os.path.exists(rep): continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    h = rr[0]
    idx = [i for i, c in enumerate(h) if c in want]
    with open(os.path.join(out, f"{tag}_{name}_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in idx]); 
        for r in rr[1:]: w.writerow([r[i][:60] for i in idx])
print("wrote", sorted(os.listdir(out)))
