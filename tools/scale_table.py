"""Markdown tables of the scaling runs: python tools/scale_table.py  (reads gpurun_out/r2_<workload>_<N>gpu.json,
copies them to profiles/)."""
import glob, json, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r2_*_?gpu.json"))):
    try:
        p = json.load(open(f))
    except Exception:
        continue
    w = os.path.basename(f)[3:].rsplit("_", 1)[0]
    rows.setdefault(w, {})[p["n_gpus"]] = p
    shutil.copy(f, os.path.join(ROOT, "profiles", os.path.basename(f)))
for w, d in rows.items():
    strong = d[min(d)]["scaling"] == "strong"
    print(f"\n**{w}** ({'strong' if strong else 'weak'} scaling)\n")
    print("| GPUs | points/s (G) | ms/step | efficiency | e2e points/s (G) | halo points sent / rank | NCCL p2p B / step / rank |")
    print("|---|---|---|---|---|---|---|")
    base = d.get(1)
    for n in sorted(d):
        p = d[n]
        v = p["value"] / 1e9
        eff = (v / (n * base["value"] / 1e9)) if base else float("nan")
        c = p.get("collectives") or {}
        hp = (c.get("halo_points") or {}).get("sent")
        e2e = ((p.get("e2e") or {}).get("value") or 0) / 1e9
        print(f"| {n} | {v:.2f} | {p['ms_per_step']:.2f} | {eff:.3f} | {e2e:.2f} | {hp} | {c.get('p2p_halo_bytes_per_step_this_rank')} |")
