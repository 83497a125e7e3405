"""One band of the C4 frame, a few frames (for an ncu launch list).  usage: _band_one.py row0 row1"""
import sys, os
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
