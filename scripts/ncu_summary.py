"""Summarise ncu reports for profiles/: (1) launch list csv (gpu__time_duration.sum) -> per-step shares,
(2) --set full report -> one table row per kernel.  usage: ncu_summary.py launches <csv> | full <report.ncu-rep>"""
import csv, subprocess, sys, collections, re


def short(name):
    m = re.search(r"(k\d\w*_kernel(?:<[^>]*>)?)", name)
    return m.group(1) if m else name[:40]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
        seq.append((short(r[ki]), v))
    # steps are delimited by k0_convert launches
    steps, cur = [], []
    for k, v in seq:
        if k.startswith("k0_convert") and cur:
            steps.append(cur); cur = []
        cur.append((k, v))
    if cur:
        steps.append(cur)
    for i, st in enumerate(steps):
        tot = sum(v for _, v in st)
        agg = collections.OrderedDict()
        for k, v in st:
            agg[k] = agg.get(k, 0) + v
        print("\n## step %d (%s frame), total %.1f us\n" % (i, "I" if not any(k.startswith("k1_") for k in agg) else "P", tot))
        print("| kernel | us | share |\n|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda x: -x[1]):
            print("| %s | %.1f | %.1f %% |" % (k, v, 100 * v / tot))


def full(path):
    # a .ncu-rep (converted here) or the raw page already exported as CSV (`ncu -i rep --page raw --csv`, done on the GPU box)
    raw = open(path, errors="ignore").read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]

    def col(r, name, default=""):
        return r[hdr.index(name)] if name in hdr else default
    print("| kernel | grid x block | time us | dram read / write MB | regs | SM throughput % | ALU pipe active % | issue active % | warps active % | smem bank conflicts |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    seen = set()
    for r in rows[2:]:
        k = short(col(r, "Kernel Name"))
        if k in seen:
            continue
        seen.add(k)

        def f(name, scale=1.0, fmt="%.1f"):
            try:
                return fmt % (float(col(r, name).replace(",", "")) * scale)
            except ValueError:
                return "-"
        tu = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(rows[1][hdr.index("gpu__time_duration.sum")], 1e-3)
        bu = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        ru, wu = bu.get(rows[1][hdr.index("dram__bytes_read.sum")], 1e-6), bu.get(rows[1][hdr.index("dram__bytes_write.sum")], 1e-6)
        print("| %s | %s x %s | %s | %s / %s | %s | %s | %s | %s | %s | %s |" % (
            k, col(r, "launch__grid_size"), col(r, "launch__block_size"), f("gpu__time_duration.sum", tu), f("dram__bytes_read.sum", ru, "%.2f"),
            f("dram__bytes_write.sum", wu, "%.2f"), col(r, "launch__registers_per_thread"), f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            f("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"), f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            f("sm__warps_active.avg.pct_of_peak_sustained_active"), f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0, "%.0f")))


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
