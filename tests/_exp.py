import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW, _lib
m = load_indexed("trex"); res=1024; V=128
f = AdvancedPixelBufferFiller(res, res, fov=45.0)
dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
views = torch.from_numpy(VW.orbit_views(V)).cuda()
z = torch.empty((V,res,res),device='cuda'); c = torch.empty((V,res,res,3),device='cuda'); n = torch.empty((V,res,res,3),device='cuda')
f.render_views(dv,dc,dn,views,z_out=z,color_out=c,normals_out=n,chunk=32)
def run(flags, label, zz=z, cc=c, nn=n):
    ptr = lambda t: None if t is None else t.data_ptr()
    def go():
        _lib.check(f._L.crb_render_views(f._handle, dv.data_ptr(), dc.data_ptr(), dn.data_ptr(), dv.shape[0], views.data_ptr(), V, ptr(zz), ptr(cc), ptr(nn), None, flags, None, f._stream()))
    for _ in range(3): go()
    f.profile(True)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): go()
    e1.record(); torch.cuda.synchronize()
    k,ms = f.profile_read(); f.profile(False)
    print(f"{label:28s} total {e0.elapsed_time(e1)/10/V*1000:6.2f} us/view   k_raster {ms/k/32*1000:6.2f} us/view")
run(0,"full")
run(0x10000,"all tiles clear-only")
run(0x20000,"no shading")
run(0x40000,"no rows (stage+shade only)")
run(0x60000,"no rows, no shading")
run(0,"full, colour only", None, c, None)
run(0,"full, z only", z, None, None)
