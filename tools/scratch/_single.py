"""Single-frame latency (one crb_render(CLEAR_FIRST) per CUDA-graph replay) with and without heavy-tile splitting."""
import sys, os, ctypes
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, _lib
for name, res in (("trex", 1024), ("bunny", 1024), ("trex", 2048)):
    m = load_indexed(name)
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
    T = dv.shape[0]
    for split in ("1", "0"):
        os.environ["CRB_SPLIT_HEAVY"] = split
        f = AdvancedPixelBufferFiller(res, res, fov=45.0)
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
        print(f"{name} {res}^2 split={split}: {a.elapsed_time(b) / 300 * 1000:7.1f} us/frame", flush=True)
