"""e2e frames/s of HostFramePipeline on the T-Rex orbit (pinned host inputs, sparse read-back).  usage: _e2e_time.py [depth] [frames]"""
import sys, os, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import views as VW
from cython3dmodelrenderer_b200.pipeline import HostFramePipeline
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 400
m = load_indexed("trex")
T = m._vertices_by_triangles.shape[0]
views = VW.orbit_views(128, 0, 7)
host_in = []
for k in range(7):
    vk, nk = VW.transform_arrays_host(views[k], m._vertices_by_triangles, m._normals_by_triangles)
    st = torch.empty((3, T, 3, 3), dtype=torch.float32).pin_memory()
    st[0].copy_(torch.from_numpy(vk)); st[1].copy_(torch.from_numpy(m._colors_by_triangles)); st[2].copy_(torch.from_numpy(nk))
    host_in.append(st)
for want in (("z", "color", "normals"), ("color",)):
    pipe = HostFramePipeline(1024, 1024, fov=45.0, depth=depth, sparse=True, want=want)
    for i in range(2 * depth):
        pipe.submit(*host_in[i % 7])
    pipe.drain(); pipe.readback_tiles()
    wait = [0.0]
    orig = pipe.result
    def timed_result(i):
        a = time.perf_counter(); r = orig(i); wait[0] += time.perf_counter() - a; return r
    pipe.result = timed_result
    t0 = time.perf_counter()
    for i in range(frames):
        pipe.submit(*host_in[i % 7])
    t1 = time.perf_counter()
    pipe.drain()
    dt = time.perf_counter() - t0
    rows = pipe.readback_tiles()
    print(f"rb_ctas={os.environ.get('CRB_READBACK_CTAS', 'default')} depth={depth} want={'+'.join(want)}: {frames / dt:8.0f} frames/s, "
          f"{rows * 1024 * sum({'z': 4, 'color': 12, 'normals': 12}[w] for w in want) / frames / 1e6:.2f} MB/frame, "
          f"{rows * 1024 * sum({'z': 4, 'color': 12, 'normals': 12}[w] for w in want) / dt / 1e9:.1f} GB/s; host: {1e6 * (t1 - t0 - wait[0]) / frames:.1f} us/frame submitting, {1e6 * wait[0] / frames:.1f} us/frame inside result()", flush=True)
    del pipe
if os.environ.get("CRB_TRACE"):
    from cython3dmodelrenderer_b200 import _lib
    _lib.load_library().crb_trace_dump(os.environ.get("CRB_TRACE_OUT", "gpurun_out/trace.txt").encode())
