#!/bin/bash
# Front-end kernel times (ncu launch lists, serialised) for the three workloads; optional CRB_LIB_OVERRIDE variant.  usage: fe_times.sh tag
tag=${1:-cur}
for w in trex sphere bunny basketball; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/l_${tag}_$w.csv python tools/scratch/prof_trex128.py $w > gpurun_out/l_${tag}_$w.log 2>&1
done
python - "$tag" <<'PY'
import csv, collections, re, sys
tag = sys.argv[1]
for w in ("trex", "sphere", "bunny", "basketball"):
    rows = [r for r in csv.reader(open(f"gpurun_out/l_{tag}_{w}.csv")) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows[len(rows) // 2:]:
        k = re.sub("<unnamed>::", "", r[4])[:14]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[14]) / 1000
    print(tag, w, "  ".join(f"{k} {us / n:.1f}" for k, (n, us) in agg.items() if k.startswith("k_")))
PY
