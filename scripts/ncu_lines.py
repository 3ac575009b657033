"""Join an ncu SASS source page (per-instruction counters) with nvdisasm line info -> per CUDA source line totals.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <object file .o> [top N]"""
import csv, subprocess, sys, re, os, tempfile, collections
rep, kre, obj = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre] + os.environ.get("NCU_LINES_ARGS", "").split(), capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
si, ii, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
kname = rows[0][1]
sass = []
for r in rows[h + 1:]:
    if len(r) <= ii or not r[0].startswith("0x"): break
    sass.append((r[si].strip(), int(r[ii] or 0), int(r[st] or 0)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# find the function section matching the kernel: pick the .text section whose instruction count matches
secs = re.split(r"\n\s*//-+ \.text\.", dis)
best = None
for s in secs[1:]:
    name = s.split()[0]
    lines = s.splitlines()
    cur = None; ins = []
    for l in lines:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m2 = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m2: ins.append((cur, m2.group(1)))
    if len(ins) == len(sass): best = (name, ins)
if not best:
    print("no section with %d instructions; sections:" % len(sass), [(s.split()[0], len(re.findall(r"/\*[0-9a-f]{4,}\*/\s+\S", s))) for s in secs[1:]]); sys.exit(1)
agg = collections.defaultdict(lambda: [0, 0])
for (loc, _), (_, n, stall) in zip(best[1], sass):
    agg[loc][0] += n; agg[loc][1] += stall
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(kname[:90]); print("section", best[0][:60], "instructions executed (warp-level):", tot, "stall samples:", tots)
srcs = {}
for loc, (n, stall) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    text = ""
    if loc:
        f = loc[0]
        if f not in srcs:
            for d in ("video-encoder_b200/csrc",):
                p = os.path.join(d, f)
                if os.path.exists(p): srcs[f] = open(p).read().splitlines()
        if f in srcs and loc[1] <= len(srcs[f]): text = srcs[f][loc[1] - 1].strip()[:90]
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * n / tot, 100 * stall / max(tots, 1), loc[0] if loc else "?", loc[1] if loc else "", text))
