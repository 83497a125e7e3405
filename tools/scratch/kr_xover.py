"""Where the small and the large rasterizer shape cross over: T-Rex x128 views at several resolutions (triangles per busy tile grow as the
frame shrinks) and UV spheres of several densities at 4096^2, each with the shape forced (1 large, 2 small) and automatic (0).
Prints the mean k_raster launch and the pairs per busy tile of the last launch."""
import os, sys, ctypes
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW, synthetic, _lib
m = load_indexed("trex"); V = 128
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
views = torch.from_numpy(VW.orbit_views(V)).cuda()
for res in (1024, 768, 512, 384, 256):
    out = []
    for shape in (1, 2, 0):
        f = AdvancedPixelBufferFiller(res, res, fov=45.0); f.set_option(_lib.CRB_OPT_RASTER_SHAPE, shape)
        z = torch.empty((V, res, res), device="cuda"); c = torch.empty((V, res, res, 3), device="cuda"); n = torch.empty((V, res, res, 3), device="cuda")
        for _ in range(3):
            f.render_views(dv, dc, dn, views, z_out=z, color_out=c, normals_out=n, chunk=V)
        torch.cuda.synchronize(); f.profile(True)
        for _ in range(10):
            f.render_views(dv, dc, dn, views, z_out=z, color_out=c, normals_out=n, chunk=V, check_status=False)
        torch.cuda.synchronize(); k, ms = f.profile_read(); f.profile(False)
        out.append(f"shape {shape}: {ms / k * 1000:7.1f} us")
        del f, z, c, n
    print(f"trex x128 {res}^2", " | ".join(out), flush=True)
for nlon, nlat in ((800, 391), (1600, 782), (2260, 1105), (3200, 1564)):
    mm = synthetic.uv_sphere(nlon, nlat); res = 4096
    sv, sc, sn = (torch.from_numpy(a).cuda() for a in (mm._vertices_by_triangles, mm._colors_by_triangles, mm._normals_by_triangles))
    out = []
    for shape in (1, 2, 0):
        f = AdvancedPixelBufferFiller(res, res, fov=45.0); f.set_option(_lib.CRB_OPT_RASTER_SHAPE, shape)
        for _ in range(3):
            f.clear(); f.render_arrays(sv, sc, sn)
        torch.cuda.synchronize(); f.profile(True)
        for _ in range(10):
            f.clear(); f.render_arrays(sv, sc, sn, check_status=False)
        torch.cuda.synchronize(); k, ms = f.profile_read(); f.profile(False)
        need, cap = ctypes.c_int64(), ctypes.c_int64()
        f._L.crb_status(f._handle, ctypes.byref(need), ctypes.byref(cap), f._stream())
        out.append(f"shape {shape}: {ms / k * 1000:7.1f} us")
        del f
    print(f"sphere {sv.shape[0]} tri {res}^2 pairs {need.value} (~{need.value / (0.7 * (res // 32) ** 2):.0f} per busy tile)", " | ".join(out), flush=True)
    del sv, sc, sn
