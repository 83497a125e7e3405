"""Where the time of the literal drop-in idiom goes (run.py:20-26: a new filler per frame, render_model, three get_*_buffer):
cProfile of 50 frames + a per-step wall-clock breakdown.  GPU box only."""
import cProfile, io, os, pstats, sys, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller
m = load_indexed("trex")


def frame():
    f = AdvancedPixelBufferFiller(1024, 1024, fov=45.0, n_threads=8)
    f.render_model(m)
    return f.get_color_buffer(), f.get_normals_buffer(), f.get_z_buffer()


for _ in range(5):
    frame()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    frame()
torch.cuda.synchronize()
print(f"new filler per frame: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms")
# step by step
acc = {}
def tick(name, t):
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t)
for _ in range(50):
    t = time.perf_counter(); f = AdvancedPixelBufferFiller(1024, 1024, fov=45.0, n_threads=8); tick("ctor", t)
    t = time.perf_counter(); f.render_model(m); tick("render_model", t)
    t = time.perf_counter(); c = f.get_color_buffer(); tick("get_color", t)
    t = time.perf_counter(); n = f.get_normals_buffer(); tick("get_normals", t)
    t = time.perf_counter(); z = f.get_z_buffer(); tick("get_z", t)
    t = time.perf_counter(); del f, c, n, z; tick("del", t)
print({k: round(v / 50 * 1e3, 3) for k, v in acc.items()})
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    frame()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
