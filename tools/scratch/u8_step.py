"""u8-only 128-view step (the payload of the row exchange) on one GPU."""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW
m = load_indexed("trex"); res, V = 1024, 128
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
views = torch.from_numpy(VW.orbit_views(V)).cuda()
u8 = torch.empty((V, res, res, 3), dtype=torch.uint8, device="cuda")
f = AdvancedPixelBufferFiller(res, res, fov=45.0)
for _ in range(3): f.render_views(dv, dc, dn, views, want=(), color_u8_out=u8, chunk=V)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f.render_views(dv, dc, dn, views, want=(), color_u8_out=u8, chunk=V, check_status=False)
e1.record(); torch.cuda.synchronize()
print("u8-only step %.1f us" % (e0.elapsed_time(e1) / 20 * 1000))
