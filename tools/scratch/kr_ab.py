"""k_raster A/B: mean launch time (CUDA events on the render stream, crb_profile) for T-Rex x128, bunny 4096^2, sphere 8192^2.
usage: [CRB_LIB_OVERRIDE=...] python tools/scratch/kr_ab.py tag"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW, synthetic
tag = sys.argv[1] if len(sys.argv) > 1 else "cur"
SHAPE = int(os.environ.get("KR_SHAPE", "0"))      # CRB_OPT_RASTER_SHAPE: 0 auto, 1 large, 2 small
from cython3dmodelrenderer_b200 import _lib
out = []
m = load_indexed("trex"); res, V = 1024, 128
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
views = torch.from_numpy(VW.orbit_views(V)).cuda()
z = torch.empty((V, res, res), device="cuda"); c = torch.empty((V, res, res, 3), device="cuda"); n = torch.empty((V, res, res, 3), device="cuda")
f = AdvancedPixelBufferFiller(res, res, fov=45.0); SHAPE and f.set_option(_lib.CRB_OPT_RASTER_SHAPE, SHAPE)
if os.environ.get('KR_CTAS'): f.set_option(_lib.CRB_OPT_RASTER_CTAS, int(os.environ['KR_CTAS']))
for _ in range(3):
    f.render_views(dv, dc, dn, views, z_out=z, color_out=c, normals_out=n, chunk=V)
torch.cuda.synchronize(); f.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    f.render_views(dv, dc, dn, views, z_out=z, color_out=c, normals_out=n, chunk=V, check_status=False)
e1.record(); torch.cuda.synchronize()
k, ms = f.profile_read(); f.profile(False)
out.append(f"trex128 k_raster {ms / k * 1000:.1f} us step {e0.elapsed_time(e1) / 20 * 1000:.1f} us")
del z, c, n, f
for name, res in (("bunny", 4096), ("sphere", 8192)):
    mm = synthetic.uv_sphere(3200, 1564) if name == "sphere" else load_indexed(name)
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (mm._vertices_by_triangles, mm._colors_by_triangles, mm._normals_by_triangles))
    f = AdvancedPixelBufferFiller(res, res, fov=45.0); SHAPE and f.set_option(_lib.CRB_OPT_RASTER_SHAPE, SHAPE)
    if os.environ.get('KR_CTAS'): f.set_option(_lib.CRB_OPT_RASTER_CTAS, int(os.environ['KR_CTAS']))
    for _ in range(3):
        f.clear(); f.render_arrays(dv, dc, dn)
    torch.cuda.synchronize(); f.profile(True)
    e0.record()
    for _ in range(20):
        f.clear(); f.render_arrays(dv, dc, dn, check_status=False)
    e1.record(); torch.cuda.synchronize()
    k, ms = f.profile_read(); f.profile(False)
    out.append(f"{name} k_raster {ms / k * 1000:.1f} us step {e0.elapsed_time(e1) / 20 * 1000:.1f} us")
    del f
print(tag, f"shape={SHAPE}", " | ".join(out))
