"""The C4 frame (10 M-triangle sphere, 8192^2) cut into N row bands, every band rendered in turn on ONE GPU (what rank r
of N does; bands are independent, so max over bands = the N-GPU frame time).  usage: _band_prof.py N [uniform|balanced]"""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, sharding, synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "balanced"
m = synthetic.uv_sphere(3200, 1564)
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
if mode == "balanced":
    bands = sharding.balanced_bands(sharding.tile_row_costs(dv, dn, 8192, 8192, 45.0), N, 8192)
else:
    bands = [sharding.band_shard(8192, r, N) for r in range(N)]
times = []
for band in bands:
    f = AdvancedPixelBufferFiller(8192, 8192, fov=45.0, band=band)
    for _ in range(3):
        f.clear(); f.render_arrays(dv, dc, dn, check_status=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f.clear(); f.render_arrays(dv, dc, dn, check_status=False)
    e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / 5)
    del f
print(f"N={N} {mode}: bands {bands}")
print(f"N={N} {mode}: ms/frame per band {[round(t, 3) for t in times]} -> max {max(times):.3f} ms = {1000 / max(times):.0f} frames/s")
