"""Where the drop-in idiom's time goes on the large configs (C2 bunny 4096^2, C4 sphere 8192^2): per-step wall clock."""
import os, sys, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import numpy as np, torch
from conftest import load_indexed, TriModel
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, synthetic
which = sys.argv[1] if len(sys.argv) > 1 else "bunny"
if which == "bunny":
    m = load_indexed("bunny"); res = 4096
else:
    m = synthetic.uv_sphere(3200, 1564); res = 8192
m = TriModel(m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles)
acc = {}
def tick(name, t):
    acc.setdefault(name, []).append(round((time.perf_counter() - t) * 1e3, 2))
keep = None
for it in range(5):
    t = time.perf_counter(); f = AdvancedPixelBufferFiller(res, res, fov=45.0, n_threads=8); tick("ctor", t)
    t = time.perf_counter(); f.render_model(m); tick("render_model", t)
    t = time.perf_counter(); torch.cuda.synchronize(); tick("sync_after_render", t)
    t = time.perf_counter(); c = f.get_color_buffer(); tick("get_color", t)
    t = time.perf_counter(); n = f.get_normals_buffer(); tick("get_normals", t)
    t = time.perf_counter(); z = f.get_z_buffer(); tick("get_z", t)
    t = time.perf_counter(); keep = (c, n, z); del f, c, n, z; tick("del", t)
for k, v in acc.items():
    print(which, k, v)
