"""Turn gpurun_out/<tag>_prof.ncu-rep (+ optional <tag>_prof_grid.ncu-rep) + <tag>_launches.csv into
profiles/<tag>_ncu_summary.md and <tag>_traffic_100M.json.
usage: python tools/make_profile_summary.py [round_tag]"""
import collections, csv, io, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u = rows[0], rows[1]
rep_grid = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_grid.ncu-rep")
if os.path.exists(rep_grid):     # the grid-ground kernels come from their own capture (same columns)
    raw_g = subprocess.run(["ncu", "-i", rep_grid, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows_g = list(csv.reader(io.StringIO(raw_g)))
    if rows_g and rows_g[0] == h:
        rows += [r for r in rows_g[2:] if "k_compact_xyz" not in "".join(r[:8]) or True]
def g(r, n): return r[h.index(n)] if n in h else ""
def f(x):
    try: return float(x.replace(",", ""))
    except Exception: return float("nan")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tsc = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
table, traffic = [], collections.defaultdict(list)
for r in rows[2:]:
    name = g(r, "Kernel Name").split("(")[0].replace("void ", "")
    dur = f(g(r, "gpu__time_duration.sum")) * tsc[u[h.index("gpu__time_duration.sum")]]
    rd = f(g(r, "dram__bytes_read.sum")) * scale[u[h.index("dram__bytes_read.sum")]]
    wr = f(g(r, "dram__bytes_write.sum")) * scale[u[h.index("dram__bytes_write.sum")]]
    table.append((name, dur * 1e3, rd / 1e9, wr / 1e9, (rd + wr) / dur / 1e9, f(g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
                  f(g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")), g(r, "launch__registers_per_thread"),
                  f(g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"))))
    if rd + wr > 1e6:                  # the surplus radix-pass launches of the device-planned sort exit at once: not a sorting pass
        traffic[name].append(rd + wr)
json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(os.path.join(ROOT, "profiles", f"{tag}_traffic_100M.json"), "w"), indent=1)
# launch list shares
lrows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv"))) if len(r) > 5]
lh = lrows[0]; ki, vi = lh.index("Kernel Name"), lh.index("Metric Value")
agg = collections.OrderedDict()
for r in lrows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except Exception: continue
    a = agg.setdefault(r[ki].split("(")[0].replace("void ", ""), [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
hot = ""
for k in ("k_pass", "k_voxel_reduce", "k_voxel_keys16", "k_db_union", "k_db_core2", "k_sum_tables"):
    hot += subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, k, "0", "7"], capture_output=True, text=True).stdout
if os.path.exists(rep_grid):
    for k in ("k_grid_min", "k_compact_xyz"):
        hot += subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep_grid, k, "0", "7"], capture_output=True, text=True).stdout
md = [f"# Round {tag[1:]} — ncu summary (B200, 100 M-point pipeline step)", "",
      "Command: `python bench.py --points 100e6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-modes` (and the same with",
      "`--ground grid` for the grid min-z kernels), run plainly first, then under",
      "`ncu --metrics gpu__time_duration.sum --clock-control none` (launch list, `" + tag + "_launches.csv`) and",
      "`ncu --set full --clock-control none --import-source on` on the heaviest kernels (second step; cold-cache and serialised:",
      "compare SHARES with the live per-kernel events in `" + tag + "_bench_1gpu.json`, not absolutes).", "",
      "## Launch list: share of the step's device time", "", "| kernel | launches | share |", "|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    md.append(f"| {k} | {a[0]} | {a[1] / tot * 100:.1f} % |")
md += ["", "## `--set full` per launch", "",
       "| kernel | ms | DRAM read GB | DRAM write GB | DRAM GB/s | DRAM % of peak | warps active % | regs | SM % |", "|---|---|---|---|---|---|---|---|---|"]
for o in table:
    md.append(f"| {o[0]} | {o[1]:.3f} | {o[2]:.3f} | {o[3]:.3f} | {o[4]:.0f} | {o[5]:.1f} | {o[6]:.1f} | {o[7]} | {o[8]:.1f} |")
md += ["", open(os.path.join(ROOT, "profiles", f"{tag}_notes.md")).read() if os.path.exists(os.path.join(ROOT, "profiles", f"{tag}_notes.md")) else "",
       "", "## Stall / hot-spot excerpts (`tools/ncu_hot.py`)", "", "```", hot, "```"]
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(md))
print("\n".join(md[:40]))
