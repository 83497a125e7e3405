import sys, time, ctypes
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, _lib
for name, res in (("trex",1024),("bunny",4096),("trex",2048)):
    m = load_indexed(name)
    f = AdvancedPixelBufferFiller(res,res,fov=45.0)
    dv,dc,dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles,m._colors_by_triangles,m._normals_by_triangles))
    for path in ("tiled","atomic"):
        for _ in range(5):
            f.clear(); f.render_arrays(dv,dc,dn,path=path)
        torch.cuda.synchronize()
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        N=200 if res<=2048 else 50
        e0.record()
        for _ in range(N):
            f._pending_clear=True
            f.render_arrays(dv,dc,dn,path=path,check_status=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/N
        T = dv.shape[0]; B = 108*T+28*res*res
        print(f"{name} {res} {path}: {ms*1000:.1f} us/frame  {1000/ms:.0f} fps  algGB/s={B/ms/1e6:.0f}")
