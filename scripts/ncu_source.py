"""Aggregates `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda [--kernel-name ...]` output: warp-level
instructions executed and stall samples per CUDA source line, top N lines by samples.
    python scripts/ncu_source.py src.csv [N]"""
import csv, sys, os
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or cur_file is None or not r[0].isdigit():
        continue
    g = lambda name: r[hdr.index(name)] if name in hdr else "0"
    try:
        samples = int(g("# Samples") or 0); inst = int(g("Instructions Executed") or 0)
    except ValueError:
        continue
    if samples or inst:
        stalls = {h[6:]: int(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h and r[i] not in ("", "0")}
        exc = g("L1 Wavefronts Shared Excessive")
        out.append((samples, inst, cur_file, int(r[0]), r[1].strip()[:90], stalls, exc))
tot_s = sum(o[0] for o in out); tot_i = sum(o[1] for o in out)
print(f"total samples {tot_s}, warp instructions {tot_i}")
for s, i, f, ln, src, st, exc in sorted(out, reverse=True)[:top]:
    top_st = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% ins  {f}:{ln:<4} {src}   [{top_st}] excess_wf={exc}")
