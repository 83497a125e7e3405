"""Is a band frame host-bound?  Host enqueue time per frame vs device time per frame.  usage: _band_host.py row0 row1"""
import sys, os, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, synthetic
band = (int(sys.argv[1]), int(sys.argv[2]))
m = synthetic.uv_sphere(3200, 1564)
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
f = AdvancedPixelBufferFiller(8192, 8192, fov=45.0, band=band)
for _ in range(4):
    f.clear(); f.render_arrays(dv, dc, dn, check_status=False)
torch.cuda.synchronize()
K = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K):
    f.clear(); f.render_arrays(dv, dc, dn, check_status=False)
t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"band {band}: host enqueue {1e3 * (t1 - t0) / K:.3f} ms/frame, device {e0.elapsed_time(e1) / K:.3f} ms/frame, wall {1e3 * (t2 - t0) / K:.3f}")
