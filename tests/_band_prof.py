"""One band of the C4 frame on one GPU (what rank r of N does): per-kernel times via CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, sharding, synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = int(sys.argv[2]) if len(sys.argv) > 2 else N // 2
m = synthetic.uv_sphere(3200, 1564)
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
band = sharding.band_shard(8192, r, N)
f = AdvancedPixelBufferFiller(8192, 8192, fov=45.0, band=band)
f.clear(); f.render_arrays(dv, dc, dn)
for _ in range(3):
    f.clear(); f.render_arrays(dv, dc, dn, check_status=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f.clear(); f.render_arrays(dv, dc, dn, check_status=False)
e1.record(); torch.cuda.synchronize()
print(f"band {band} of N={N}: {e0.elapsed_time(e1) / 5:.3f} ms/frame")
