#!/usr/bin/env python
"""Where warps wait: for one launch of an ncu report, the SASS instructions with the most stall samples of a given reason,
with the source line they belong to.  usage: tools/ncu_stalls.py report.ncu-rep [reason=barrier] [top=15] [launch=0]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; reason = sys.argv[2] if len(sys.argv) > 2 else "barrier"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# blocks per file; the first file of each launch is the .cu
launches, cur, first_file, cur_file, hdr = [], None, None, None, None
for row in rows:
    if not row: continue
    if row[0] == "File Path":
        cur_file = row[1]
        if first_file is None: first_file = cur_file
        if cur_file == first_file: cur = []; launches.append(cur)
        continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if cur is not None: cur.append((cur_file, row))
col = hdr.index("stall_" + reason)
tot_col = hdr.index("# Samples")
seen = set(); items = []; line = None; total = 0; total_all = 0
for f, r in launches[which]:
    if r[0].isdigit():
        line = (f.split("/")[-1], int(r[0]), r[1].strip()); continue
    if r[0] == "" and r[2].startswith("0x"):
        if r[2] in seen: continue
        seen.add(r[2])
        try: n = int(r[col]); a = int(r[tot_col])
        except ValueError: continue
        total += n; total_all += a
        if n: items.append((n, line, r[3].strip()))
items.sort(key=lambda t: -t[0])
print(f"stall_{reason}: {total} samples of {total_all} ({100.0*total/max(total_all,1):.1f} %)")
for n, ln, sass in items[:top]:
    print(f"{n:7d} {100.0*n/max(total,1):5.1f}%  {ln[0][:14]}:{ln[1]:<5d} {sass[:40]:40s} | {ln[2][:70]}")
