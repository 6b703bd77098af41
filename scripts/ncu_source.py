"""Aggregates `ncu --page source --csv --print-source sass,cuda` output: instructions executed and
stall samples per CUDA source line (file:line), for the first kernel instance in the file."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# find header rows ("Line No"/"Address" variants); SASS view rows carry the source location in a column
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] in ("Line No", "Address", "#")]
print("sections:", len(hdr_idx), [rows[i][:4] for i in hdr_idx[:4]])
