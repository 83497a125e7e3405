"""Frame time of single bands of the C4 frame (one GPU; bands are independent, the slowest band is the N-GPU frame time).
usage: band_time.py row0:row1 [row0:row1 ...]"""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, synthetic
m = synthetic.uv_sphere(3200, 1564)
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
out = []
for spec in sys.argv[1:]:
    r0, r1 = (int(x) for x in spec.split(":"))
    f = AdvancedPixelBufferFiller(8192, 8192, fov=45.0, band=(r0, r1))
    for _ in range(3):
        f.clear(); f.render_arrays(dv, dc, dn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f.clear(); f.render_arrays(dv, dc, dn, check_status=False)
    e1.record(); torch.cuda.synchronize()
    out.append(f"{spec} {e0.elapsed_time(e1) / 10 * 1000:.0f} us")
    del f
print(" | ".join(out))
