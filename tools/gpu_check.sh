#!/bin/bash
# Runs on the GPU box: parity tests, quick per-config timing, bench summary.  usage: tools/gpu_check.sh [pytest-args]
timeout 900 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -4
python tools/scratch/_quick_time.py 2>&1 | grep -v Warning
timeout 600 python bench.py --steps 10 --no-cpu 2>/dev/null > gpurun_out/bench_quick.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
r = d["roofline"]
print(f"value={d['value']:.0f} fps  us/view={1e6/d['value']:.2f}  k_raster frac={r['frac']:.3f} avg_launch_ms={r['avg_launch_ms']:.3f} share={r['share_of_step']:.2f}")
print("single_frame", d["single_frame"].get("us_per_frame"), "e2e", d["e2e"]["value"], "clocks", d["clocks"])
PY
