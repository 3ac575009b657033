"""per-phase shares of one kernel from an `ncu --page source --csv` export: the kernel's SASS is cut at its CTA barriers / mbarrier
waits and, per phase, the share of executed warp-instructions, of warp-stall samples and the shared-memory wavefronts (total,
excess = bank conflicts) are printed.  usage: ncu_phase_shares.py source.csv [table-index]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r)]
ti = int(sys.argv[2]) if len(sys.argv) > 2 else len(starts) - 1
st = starts[ti]; h = rows[st]; idx = {c: i for i, c in enumerate(h)}
end = starts[ti + 1] - 1 if ti + 1 < len(starts) else len(rows)
body = [r for r in rows[st + 1:end] if len(r) == len(h)]
f = lambda r, c: float(r[idx[c]] or 0)
regs = []; cur = dict(name="start", inst=0, samp=0, exc=0, wf=0, n=0)
for r in body:
    src = r[idx["Source"]]
    cur["inst"] += f(r, "Instructions Executed"); cur["samp"] += f(r, "# Samples")
    cur["exc"] += f(r, "L1 Wavefronts Shared Excessive"); cur["wf"] += f(r, "L1 Wavefronts Shared"); cur["n"] += 1
    if "BAR.SYNC" in src or "SYNCS.PHASECHK" in src or "SYNCS.ARRIVE" in src:
        regs.append(cur); cur = dict(name="after " + src.strip()[:34], inst=0, samp=0, exc=0, wf=0, n=0)
regs.append(cur)
ti_ = sum(x["inst"] for x in regs) or 1; ts_ = sum(x["samp"] for x in regs) or 1
print("table %d of %d: %d SASS lines, %.1f M warp-instructions, %d stall samples" % (ti, len(starts), len(body), ti_ / 1e6, ts_))
for x in regs:
    print("%-42s SASS lines %4d  instructions %5.1f%%  stall samples %5.1f%%  smem wavefronts %10.0f  excess %10.0f"
          % (x["name"], x["n"], 100 * x["inst"] / ti_, 100 * x["samp"] / ts_, x["wf"], x["exc"]))
