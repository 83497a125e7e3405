"""Small driver for ncu: a few single T-Rex frames (tiled + atomic path) and one 32-view batched launch."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW
which = sys.argv[1] if len(sys.argv) > 1 else "trex"
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nviews = int(sys.argv[3]) if len(sys.argv) > 3 else 32
m = load_indexed(which)
f = AdvancedPixelBufferFiller(res, res, fov=45.0)
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
for i in range(4):
    f.clear(); f.render_arrays(dv, dc, dn)
for i in range(2):
    f.clear(); f.render_arrays(dv, dc, dn, path="atomic")
if nviews:
    views = VW.orbit_views(128, 0, nviews)
    out = None
    for i in range(3):
        out = f.render_views(dv, dc, dn, views, chunk=nviews, z_out=None if out is None else out["z"],
                             color_out=None if out is None else out["color"], normals_out=None if out is None else out["normals"])
torch.cuda.synchronize()
print("done", f.launch_count)
