"""A/B timing of k_raster variants selected by environment switches read at filler creation (GPU box only)."""
import sys, os, ctypes
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
if any("CRB_DEBUG_SKIP" in a for a in sys.argv[1:]) or True:
    _abl = os.path.join(_ROOT, "cython3dmodelrenderer_b200", "csrc", "libcrender_b200_ablation.so")
    if os.path.exists(_abl) and "CRB_LIB_OVERRIDE" not in os.environ:
        os.environ["CRB_LIB_OVERRIDE"] = _abl      # the product build has no ablation switches
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW, _lib
m = load_indexed("trex"); res = 1024; V = 128
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
views = torch.from_numpy(VW.orbit_views(V)).cuda()
z = torch.empty((V, res, res), device='cuda'); c = torch.empty((V, res, res, 3), device='cuda'); n = torch.empty((V, res, res, 3), device='cuda')


def run(label, env, zz=z, cc=c, nn=n):
    for k in ("CRB_NO_TMA", "CRB_RASTER_CTAS", "CRB_OUT_TMA", "CRB_DEBUG_SKIP", "CRB_TILES_PER_CTA"):
        os.environ.pop(k, None)
    os.environ.update(env)
    f = AdvancedPixelBufferFiller(res, res, fov=45.0)
    f.render_views(dv, dc, dn, views, z_out=zz, color_out=cc, normals_out=nn, chunk=32)
    ptr = lambda t: None if t is None else t.data_ptr()

    def go():
        _lib.check(f._L.crb_render_views(f._handle, dv.data_ptr(), dc.data_ptr(), dn.data_ptr(), dv.shape[0], views.data_ptr(), V,
                                         ptr(zz), ptr(cc), ptr(nn), None, 0, None, f._stream()))
    for _ in range(3):
        go()
    f.profile(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record(); torch.cuda.synchronize()
    k, ms = f.profile_read(); f.profile(False)
    print(f"{label:34s} total {e0.elapsed_time(e1)/10/V*1000:6.2f} us/view   k_raster {ms/k/32*1000:6.2f} us/view", flush=True)


run("default", {})
run("shade from record 0 (no gathers)", {"CRB_DEBUG_SKIP": "16"})
run("no shading", {"CRB_DEBUG_SKIP": "2"})
run("no rows", {"CRB_DEBUG_SKIP": "4"})
