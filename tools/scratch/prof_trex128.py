"""Driver for an ncu capture of the headline k_raster launch: T-Rex 1024^2, 128 orbit views in one launch, three launches.
ncu -k regex:k_raster --launch-skip 2 --launch-count 1 --set full --import-source on --clock-control none -o rep python tools/scratch/prof_trex128.py [workload]"""
import os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW, synthetic
wl = sys.argv[1] if len(sys.argv) > 1 else "trex"
if wl == "trex":
    m = load_indexed("trex"); res, V = 1024, 128
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
    views = torch.from_numpy(VW.orbit_views(V)).cuda()
    z = torch.empty((V, res, res), device="cuda"); c = torch.empty((V, res, res, 3), device="cuda"); n = torch.empty((V, res, res, 3), device="cuda")
    f = AdvancedPixelBufferFiller(res, res, fov=45.0)
    for _ in range(3):
        f.render_views(dv, dc, dn, views, z_out=z, color_out=c, normals_out=n, chunk=V)
else:
    if wl == "sphere":
        m = synthetic.uv_sphere(3200, 1564); res = 8192
    elif wl == "basketball":
        m = load_indexed("basketball"); res = 2048
    else:
        m = load_indexed("bunny"); res = 4096
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
    f = AdvancedPixelBufferFiller(res, res, fov=45.0)
    for _ in range(3):
        f.clear(); f.render_arrays(dv, dc, dn)
torch.cuda.synchronize()
