"""Single-frame latency (one crb_render(CLEAR_FIRST) per CUDA-graph replay) per rasterizer shape (CRB_OPT_RASTER_SHAPE 0 auto / 1 large / 2 small)."""
import sys, os, ctypes
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, _lib
tag = sys.argv[1] if len(sys.argv) > 1 else "cur"
shapes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2").split(",")]
for name, res in (("trex", 1024), ("bunny", 1024), ("trex", 2048), ("bunny", 4096)):
    m = load_indexed(name)
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
    T = dv.shape[0]
    out = []
    for shape in shapes:
        f = AdvancedPixelBufferFiller(res, res, fov=45.0)
        if shape:
            f.set_option(_lib.CRB_OPT_RASTER_SHAPE, shape)
        f.clear(); f.render_arrays(dv, dc, dn)
        L, h = f._L, f._handle
        g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                _lib.check(L.crb_render(h, dv.data_ptr(), dc.data_ptr(), dn.data_ptr(), T, _lib.CRB_CLEAR_FIRST, f._stream()))
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                _lib.check(L.crb_render(h, dv.data_ptr(), dc.data_ptr(), dn.data_ptr(), T, _lib.CRB_CLEAR_FIRST, f._stream()))
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(300):
                g.replay()
            b.record(); torch.cuda.synchronize()
        out.append(f"shape {shape}: {a.elapsed_time(b) / 300 * 1000:7.1f} us")
    print(tag, f"{name} {res}^2", " | ".join(out), flush=True)
