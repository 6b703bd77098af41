"""Splits the SASS listing of the first kernel in an `ncu --page source --csv --print-source sass`
dump into regions delimited by BAR.SYNC and prints instructions executed / stall samples per
region plus the opcode mix."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hi[0]
end = hi[1] - 2 if len(hi) > 1 else len(rows)
hdr = rows[start]
ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "# Samples") if n in hdr}
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
regions, cur = [], dict(n=0, inst=0, samp=0, ops=collections.Counter(), stalls=collections.Counter(), first=None)
total = 0
for r in rows[start + 1:end]:
    if len(r) < len(hdr): continue
    src = r[ci["Source"]].strip()
    try: ie = int(r[ci["Instructions Executed"]])
    except ValueError: continue
    sm = int(r[ci["# Samples"]]) if r[ci["# Samples"]].isdigit() else 0
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    cur["n"] += 1; cur["inst"] += ie; cur["samp"] += sm; cur["ops"][op] += ie
    for i, h in stall_cols:
        if r[i].isdigit(): cur["stalls"][h] += int(r[i])
    total += ie
    if cur["first"] is None: cur["first"] = r[0]
    if src.startswith("BAR.SYNC") or "BAR.SYNC" in src:
        regions.append(cur)
        cur = dict(n=0, inst=0, samp=0, ops=collections.Counter(), stalls=collections.Counter(), first=None)
regions.append(cur)
print("total warp-instructions:", total)
for k, g in enumerate(regions):
    if g["inst"] == 0: continue
    top = ", ".join(f"{o}:{c*100//max(1,g['inst'])}%" for o, c in g["ops"].most_common(8))
    st = ", ".join(f"{h[6:]}:{c}" for h, c in g["stalls"].most_common(5))
    print(f"region {k:2d} @{g['first']} sass={g['n']:5d} inst={g['inst']:10d} ({100*g['inst']/total:5.1f}%) samples={g['samp']:6d} | {top} | {st}")
