"""Regenerates profiles/ncu_traffic.json from an ncu launch list of the dominant kernel, stamped with the hash of the
kernel's sources (bench.source_hash): run it in the same commit as any change to csrc/jrc_{common,exact,fused}.cuh.
    gpurun -- 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
               -k regex:k_fused64x8 -s 4 -c 3 --csv --log-file gpurun_out/fused_traffic.csv \\
               python bench.py --steps 3 --warmup 3 --no-configs --no-cpu'
    python scripts/update_traffic.py gpurun_out/fused_traffic.csv"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
launches = {}
for r in rows[1:]:
    if "k_fused64x8<16, 8, 0, 1>" in r[ki]:
        launches.setdefault(r[ii], {})[r[mi]] = int(float(r[vi].replace(",", "")))
if not launches:
    raise SystemExit("no k_fused64x8<16, 8, 0, 1> launch in " + sys.argv[1])
ids = sorted(launches, key=int)
first = launches[ids[0]]
totals = [launches[i]["dram__bytes_read.sum"] + launches[i]["dram__bytes_write.sum"] for i in ids]
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
j = json.load(open(path)) if os.path.exists(path) else {}
j["k_fused64x8"] = {
    "dram_bytes_per_launch": totals[0], "dram_bytes_read": first["dram__bytes_read.sum"], "dram_bytes_write": first["dram__bytes_write.sum"],
    "source_hash": bench.source_hash(),
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, k_fused64x8<16,8,false,true> with the "
              "in-kernel estimator, 4096 CPIs per launch, first of %d captured launches of `bench.py --steps 3 --warmup 3` "
              "(all: %s B)" % (len(ids), ", ".join(str(t) for t in totals)),
    "note": "below the algorithmic 1124204544 B because ~60 MB of map lines are still dirty in the 126 MB L2 when the kernel ends; no re-reads",
}
json.dump(j, open(path, "w"), indent=1)
print(json.dumps(j["k_fused64x8"], indent=1))
