"""Turns one round's gpurun_out captures into the tracked summaries under profiles/.

usage: python scripts/summarise_profiles.py TAG launches.csv prof.ncu-rep bench.json EVALS_PER_LAUNCH
"""
import collections
import csv
import json
import shutil
import subprocess
import sys

tag, launches, rep, bench, evals = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
rows = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
lines = [f"# ncu launch list ({tag}): cold-cache, serialised -- compare SHARES, not absolutes",
         "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv python bench.py --steps 3 --warmup 3 --no-cpu",
         "# (covers image creation, both k-means inits, warm-up and timed steps of the device-resident and host-buffer arms)",
         f"{'kernel':44s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k[:44]:44s} {n:8d} {us / 1e3:10.3f} {us / n:10.1f} {100 * us / tot:6.1f}%")
open(f"profiles/{tag}_launches_summary.txt", "w").write("\n".join(lines) + "\n")
shutil.copy(launches, f"profiles/{tag}_launches.csv")
shutil.copy(bench, f"profiles/{tag}_bench_n1.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(txt.splitlines()))
h, units = rr[0], rr[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
keys += [x for x in h if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio")]
out = []
for r in rr[2:]:
    d = {"kernel": r[h.index("Kernel Name")]}
    for k in keys:
        if k in h:
            d[k] = {"value": r[h.index(k)], "unit": units[h.index(k)]}
    out.append(d)
json.dump(out, open(f"profiles/{tag}_ncu_k_score_v3_raw_subset.json", "w"), indent=1)
d = max(out, key=lambda r: float(r["smsp__inst_executed.sum"]["value"]))   # the per-kernel figures quoted below: the larger of the step's scorer kernels


def nbytes(x):
    return float(x["value"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[x["unit"]]


sys.path.insert(0, ".")
from snesimage_b200 import _build  # noqa: E402

rd = sum(nbytes(r["dram__bytes_read.sum"]) for r in out)
wr = sum(nbytes(r["dram__bytes_write.sum"]) for r in out)
j = {"kernel": "k_score_pair+k_score_v3" if len(out) > 1 else "k_score_v3", "scorer_source_sha256": _build.scorer_source_hash(),
     "dram_bytes_per_eval": (rd + wr) / evals,
     "source": f"ncu --set full --clock-control none, the scorer kernels of one step ({', '.join(r['kernel'].split('(')[0].split('::')[-1] for r in out)}) over {evals} evaluations, "
               f"profiles/{tag}_ncu_k_score_v3_raw_subset.json: dram__bytes_read.sum {rd / 1e9:.4f} GB + dram__bytes_write.sum {wr / 1e6:.3f} MB",
     "inst_executed_per_eval": sum(float(r["smsp__inst_executed.sum"]["value"]) for r in out) / evals,
     "ipc_active": float(d["sm__inst_executed.avg.per_cycle_active"]["value"]),
     "issue_slots_busy_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"]),
     "l2_hit_pct": float(d["lts__t_sector_hit_rate.pct"]["value"])}
json.dump(j, open("profiles/roofline_traffic.json", "w"), indent=1)
print("\n".join(lines[3:12]))
print(json.dumps(j, indent=1))
print({k.split("stalled_")[1].split("_per")[0]: round(float(v["value"]), 2) for k, v in d.items() if "stalled" in k})
print({k.split(".")[0]: d[k]["value"] for k in d if "pipe" in k}, d["gpu__time_duration.sum"])
