"""profiles/sass_opcodes.txt: per-kernel histogram of the Blackwell-specific SASS opcodes in libdp_b200.so
(`cuobjdump -sass`): UTCHMMA/UTCQMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), UTCATOMSWS (TMEM alloc), plus HMMA / IMMA (mma.sync, must be absent)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "disruption-prediciton-based-on-multimodal-deep-learning_b200", "libdp_b200.so")
ops = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UTMAPF", "UTMACCTL", "UTMACMDFLUSH", "HMMA", "IMMA", "FFMA"]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
hist = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        hist[cur]["_total"] += 1
        for o in ops:
            if op.startswith(o):
                hist[cur][o] += 1
out = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
tot = collections.Counter()
with open(out, "w") as f:
    f.write("cuobjdump -sass libdp_b200.so (sm_100a): SASS opcode counts per kernel (scripts/sass_opcodes.py)\n")
    f.write("UTCHMMA = tcgen05.mma kind::f16, UTMALDG/UTMASTG = TMA bulk tensor load/store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,\n"
            "SYNCS = mbarrier ops, UTCATOMSWS = tcgen05.alloc/dealloc; HMMA/IMMA = legacy mma.sync (none expected)\n\n")
    f.write(f"{'kernel':80s} " + " ".join(f"{o:>9s}" for o in ops) + "     total\n")
    for k, c in hist.items():
        for o in ops:
            tot[o] += c[o]
        if any(c[o] for o in ops[:10]):
            f.write(f"{k[-80:]:80s} " + " ".join(f"{c[o]:9d}" for o in ops) + f" {c['_total']:9d}\n")
    f.write(f"\n{'ALL KERNELS (' + str(len(hist)) + ')':80s} " + " ".join(f"{tot[o]:9d}" for o in ops) + "\n")
print(open(out).read()[-1500:])
