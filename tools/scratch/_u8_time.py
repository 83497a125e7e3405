"""uint8-image-only batched render (N3 fused into the rasterizer): 128 T-Rex views, chunks of 32 / 128."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW
m = load_indexed("trex")
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
V = 128
views = torch.from_numpy(VW.orbit_views(128, 0, V)).cuda()
f = AdvancedPixelBufferFiller(1024, 1024, fov=45.0)
u8 = torch.empty((V, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
for chunk in (32, 128):
    go = lambda: f.render_views(dv, dc, dn, views, want=(), color_u8_out=u8, chunk=chunk, check_status=False)
    f.render_views(dv, dc, dn, views, want=(), color_u8_out=u8, chunk=chunk)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record(); torch.cuda.synchronize()
    print(f"u8 only, chunk {chunk}: {e0.elapsed_time(e1) / 10:.3f} ms per {V} views = {V * 10 / e0.elapsed_time(e1) * 1000:.0f} images/s; lit pixels view 0: {int((u8[0].sum(dim=-1) > 0).sum())}", flush=True)
