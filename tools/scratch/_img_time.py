"""HostImagePipeline rate and where the host time goes.  usage: _img_time.py [depth] [frames]"""
import sys, os, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import numpy as np, torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import views as VW, HostImagePipeline
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 400
m = load_indexed("trex")
views = VW.orbit_views(128, 0, 5)
host_in = []
for k in range(5):
    vk, nk = VW.transform_arrays_host(views[k], m._vertices_by_triangles, m._normals_by_triangles)
    host_in.append(torch.from_numpy(np.stack([vk, m._colors_by_triangles, nk])).pin_memory())
pipe = HostImagePipeline(1024, 1024, fov=45.0, depth=depth)
for i in range(2 * depth):
    pipe.submit(host_in[i % 5])
pipe.drain()
wait = [0.0]
orig = pipe.result
def timed(i):
    a = time.perf_counter(); r = orig(i); wait[0] += time.perf_counter() - a; return r
pipe.result = timed
t0 = time.perf_counter()
for i in range(frames):
    pipe.submit(host_in[i % 5])
t1 = time.perf_counter()
pipe.drain()
dt = time.perf_counter() - t0
print(f"depth={depth}: {frames / dt:8.0f} images/s; host {1e6 * (t1 - t0 - wait[0]) / frames:.1f} us/frame submitting, {1e6 * wait[0] / frames:.1f} us/frame inside result()")
