"""Per-kernel summary of the `ncu --set full` stage captures (scripts/stage_capture.sh): one line per kernel with the launch that
ran longest -- duration, DRAM and L2 bytes and rates, SM / memory throughput, occupancy.

usage: python scripts/summarise_stages.py TAG mode=report.ncu-rep|raw.csv [...]   -> profiles/TAG_stages.txt
"""
import csv
import subprocess
import sys

tag = sys.argv[1]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
out = [f"# {tag}: ncu --set full --clock-control none, one launch per kernel (the longest of those captured), scripts/quick_bench.py 4 <mode> v3",
       "# (4 images x 64 candidates = 256 evaluations per launch; DRAM peak of this pool 6552.6 GB/s measured, MEASURED_PEAKS.json)",
       f"{'mode':7s}{'kernel':34s}{'grid':>8s}{'ms':>9s}{'DRAM MB':>10s}{'DRAM GB/s':>11s}{'L2 MB':>10s}{'L2 GB/s':>10s}{'dram%':>7s}{'sm%':>6s}{'occ%':>6s}{'regs':>6s}{'IPC':>6s}"]
for spec in sys.argv[2:]:
    mode, rep = spec.split("=")
    # a .ncu-rep, or the raw-page CSV made from it on the GPU box (the reports are too big to bring back)
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        out.append(f"{mode}: empty report")
        continue
    h, units = rows[0], rows[1]

    def col(r, name, scale=None):
        if name not in h:
            return float("nan")
        i = h.index(name)
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return float("nan")
        if scale is not None:
            v *= scale.get(units[i], 1.0)
        return v

    best = {}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        short = name.split("snes::")[-1].split("(")[0]
        t = col(r, "gpu__time_duration.sum", TIME)
        if short not in best or t > best[short][0]:
            best[short] = (t, r)
    for short, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
        dram = col(r, "dram__bytes_read.sum", UNIT) + col(r, "dram__bytes_write.sum", UNIT)
        l2 = col(r, "lts__t_bytes.sum", UNIT)
        if l2 != l2:   # this ncu version reports L2 traffic in 32-byte sectors
            l2 = col(r, "lts__t_sectors.sum") * 32.0
        out.append(f"{mode:7s}{short[:33]:34s}{int(col(r, 'launch__grid_size')):8d}{t * 1e3:9.3f}{dram / 1e6:10.1f}{dram / t / 1e9:11.1f}{l2 / 1e6:10.1f}{l2 / t / 1e9:10.1f}"
                   f"{col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):7.1f}{col(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f}"
                   f"{col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f}{int(col(r, 'launch__registers_per_thread')):6d}{col(r, 'sm__inst_executed.avg.per_cycle_active'):6.2f}")
open(f"profiles/{tag}_stages.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
