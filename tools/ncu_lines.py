#!/usr/bin/env python
"""Per-CUDA-source-line summary of an ncu report (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py report.ncu-rep [launch_index=0] [top_n=40]
Prints, for the chosen profiled launch, the source lines ranked by warp-instructions executed, with their share of
all instructions and of all warp-stall samples."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
launches, cur, first_file, cur_file = [], None, None, None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path":
        cur_file = row[1]
        if first_file is None: first_file = cur_file
        if cur_file == first_file:
            cur = []; launches.append(cur)
        continue
    if row[0] == "Function Name":
        cur.append({"name": row[1], "file": cur_file, "rows": []}); continue
    if row[0] == "Line No" or cur is None: continue
    if row[0].isdigit() and cur_file == first_file: cur[-1]["rows"].append(row)
L = launches[which]
inst = collections.Counter(); samp = collections.Counter(); src = {}
for b in L:
    for r in b["rows"]:
        ln = int(r[0]); src[ln] = r[1]
        try: inst[ln] += int(r[7]); samp[ln] += int(r[6])
        except ValueError: pass
tot, totS = sum(inst.values()), sum(samp.values())
print(f"launch {which}/{len(launches)}: {L[0]['name']}  warp-inst={tot}  samples={totS}")
for ln, n in inst.most_common(top):
    print(f"{ln:>5} {n:>10} {100*n/max(tot,1):5.1f}%  st={100*samp[ln]/max(totS,1):4.1f}%  {src[ln].strip()[:105]}")
