"""Per-config frame time (device-resident, fused clear): T-Rex 1024 batch of 32 views, bunny 2048 / 4096 single frames."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW
for name, res, V in (("trex", 1024, 32), ("bunny", 2048, 1), ("bunny", 4096, 1), ("trex", 4096, 1)):
    m = load_indexed(name)
    f = AdvancedPixelBufferFiller(res, res, fov=45.0)
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
    views = torch.from_numpy(VW.orbit_views(128, 0, V)).cuda()
    out = f.render_views(dv, dc, dn, views, chunk=V)
    go = lambda: f.render_views(dv, dc, dn, views, chunk=V, z_out=out["z"], color_out=out["color"], normals_out=out["normals"], check_status=False)
    for _ in range(3):
        go()
    f.profile(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 20
    e0.record()
    for _ in range(N):
        go()
    e1.record(); torch.cuda.synchronize()
    k, ms = f.profile_read(); f.profile(False)
    B = 108 * dv.shape[0] + 28 * res * res
    print(f"{name} {res}^2 x{V}: {e0.elapsed_time(e1)/N/V*1000:8.1f} us/frame   k_raster {ms/k/V*1000:8.1f} us/frame  = {B/(ms/k/V)/1e6:6.0f} GB/s", flush=True)
