#!/usr/bin/env python
"""Sum ncu per-line instruction counts over named source line ranges.  usage: ncu_ranges.py report launch name:lo-hi[,lo-hi] ..."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]); specs = sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
launches, cur, first_file, cur_file = [], None, None, None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path":
        cur_file = row[1]
        if first_file is None: first_file = cur_file
        if cur_file == first_file: cur = collections.Counter(); thr = collections.Counter(); launches.append((cur, thr))
        continue
    if row[0] in ("Function Name", "Line No") or cur is None: continue
    if row[0].isdigit() and cur_file == first_file:
        try: cur[int(row[0])] += int(row[7]); thr[int(row[0])] += int(row[8])
        except ValueError: pass
inst, thr = launches[which]; tot = sum(inst.values()); acc = 0
for sp in specs:
    name, rngs = sp.split(":"); n = t = 0
    for rg in rngs.split(","):
        lo, hi = map(int, rg.split("-"))
        n += sum(v for k, v in inst.items() if lo <= k <= hi); t += sum(v for k, v in thr.items() if lo <= k <= hi)
    acc += n
    print(f"{name:24s} {n:>11} {100*n/tot:5.1f}%   lanes/inst {t/max(n,1):5.1f}")
print(f"{'(other lines)':24s} {tot-acc:>11} {100*(tot-acc)/tot:5.1f}%   total {tot}")
