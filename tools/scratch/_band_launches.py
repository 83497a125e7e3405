"""Mean duration per kernel name from an ncu launch list.  usage: _band_launches.py file.csv [skip_first_n]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = collections.defaultdict(list)
for r in rows[1 + skip:]:
    d[r[ki].split("(")[0].split("::")[-1]].append(float(r[vi].replace(",", "")))
for k, v in d.items():
    print(f"{k:28s} n={len(v):3d} mean {sum(v) / len(v) / 1000:9.2f} us  min {min(v) / 1000:9.2f} us")
