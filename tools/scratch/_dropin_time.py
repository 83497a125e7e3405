"""Drop-in flow timing: what a reference user sees per frame (host NumPy model in, host NumPy buffers out)."""
import sys, os, time
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import numpy as np, torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller
m = load_indexed("trex")
def t(fn, n=30, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def new_filler_per_frame():
    f = AdvancedPixelBufferFiller(1024, 1024, fov=45.0, n_threads=8)
    f.render_model(m)
    return f.get_color_buffer(), f.get_normals_buffer(), f.get_z_buffer()
f = AdvancedPixelBufferFiller(1024, 1024, fov=45.0)
def reuse_clear():
    f.clear(); f.render_model(m)
    return f.get_color_buffer(), f.get_normals_buffer(), f.get_z_buffer()
def reuse_clear_color_only():
    f.clear(); f.render_model(m)
    return f.get_color_buffer()
def render_only():
    f.clear(); f.render_model(m)
print(f"new filler per frame (run.py idiom), all 3 buffers : {t(new_filler_per_frame):7.3f} ms/frame")
print(f"one filler, clear() + render_model + 3 buffers       : {t(reuse_clear):7.3f} ms/frame")
print(f"one filler, clear() + render_model + colour only     : {t(reuse_clear_color_only):7.3f} ms/frame")
print(f"one filler, clear() + render_model (no read-back)    : {t(render_only):7.3f} ms/frame")
