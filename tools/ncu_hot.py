"""Summarise an ncu source page: top SASS lines by stall samples + stall reason totals.
usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [launch_skip]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci, cs = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = 0; agg = []; st = {h: 0 for _, h in stall_cols}
for r in rows[2:]:
    try: n = float(r[cs])
    except Exception: continue
    tot += n; agg.append((n, r[ci].strip()))
    for i, h in stall_cols:
        try: st[h] += float(r[i])
        except Exception: pass
print(rows[0][1][:100], "samples", tot)
print("stalls:", ", ".join(f"{h[6:]}={v/ max(1,sum(st.values()))*100:.0f}%" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
for n, sline in sorted(agg, reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 18]:
    print(f"{n/tot*100:5.1f}%  {sline[:100]}")
