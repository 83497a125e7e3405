"""ctypes front-end of the CPU oracle (oracle/crender_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports it and has no CPU fallback.

`OracleFiller` mirrors the reference class
(crender/cy/pixel_buffer_filler/advanced_pixel_buffer_filler.pyx:20-253): same constructor, same
`render_model(model)` duck-typing on `_vertices_by_triangles/_colors_by_triangles/_normals_by_triangles`,
same persistent buffers and live views.  Parity status: pinned (see the header of crender_oracle.c).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile liboracle.so with gcc (seconds).  Building the checker is not using it."""
    newest = max(os.path.getmtime(os.path.join(_HERE, s)) for s in ("crender_oracle.c", "ingest_oracle.c"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        L.oracle_projection.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_float, fp]
        L.oracle_projection.restype = ctypes.c_int
        L.oracle_project_vertex.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp]
        L.oracle_project_vertex.restype = None
        L.oracle_pixel_rect.argtypes = [fp, ctypes.c_int, ctypes.c_int, ip]
        L.oracle_pixel_rect.restype = None
        L.oracle_barycentric.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp]
        L.oracle_barycentric.restype = None
        L.oracle_render.argtypes = [ctypes.c_int, ctypes.c_int, fp, fp, fp, fp, ctypes.c_int64, fp, fp, fp,
                                    ctypes.c_int]
        L.oracle_render.restype = ctypes.c_int
        L.oracle_project.argtypes = [ctypes.c_int, ctypes.c_int, fp, fp, ctypes.c_int64, fp]
        L.oracle_project.restype = None
        L.oracle_init_buffers.argtypes = [ctypes.c_int, ctypes.c_int, fp, fp, fp]
        L.oracle_init_buffers.restype = None
        L.oracle_guro.argtypes = [ctypes.c_int, ctypes.c_int, fp, fp, fp]
        L.oracle_guro.restype = None
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _f32_tri_array(a, name):
    a = np.asarray(a)
    if a.dtype != np.float32:
        # the reference's typed memoryviews raise exactly this
        raise ValueError(f"Buffer dtype mismatch, expected 'float' but got '{a.dtype.name}' ({name})")
    if a.ndim != 3 or a.shape[1:] != (3, 3):
        raise ValueError(f"{name}: expected shape [T,3,3], got {a.shape}")
    return np.ascontiguousarray(a)


def projection_matrix(h, w, fov=90.0, z_near=0.1, z_far=1000.0):
    P = np.zeros(16, dtype=np.float32)
    if lib().oracle_projection(int(h), int(w), float(fov), float(z_near), float(z_far), _fp(P)) != 0:
        raise ZeroDivisionError("float division")
    return P.reshape(4, 4)


def project(h, w, P, v):
    v = _f32_tri_array(v, "vertices")
    out = np.empty_like(v)
    lib().oracle_project(int(h), int(w), _fp(np.ascontiguousarray(P, dtype=np.float32).ravel()),
                         _fp(v), v.shape[0], _fp(out))
    return out


def pixel_rect(tri, h, w):
    tri = np.ascontiguousarray(tri, dtype=np.float32).ravel()
    out = np.zeros(4, dtype=np.int32)
    lib().oracle_pixel_rect(_fp(tri), int(h), int(w), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return tuple(int(x) for x in out)


def barycentric(tri, x, y):
    tri = np.ascontiguousarray(tri, dtype=np.float32).ravel()
    out = np.zeros(3, dtype=np.float32)
    lib().oracle_barycentric(_fp(tri), int(x), int(y), _fp(out))
    return out


def guro(color, normals, light_direction=(0, 0, 1)):
    """In-place Guro illumination of `color` ([h,w,3] f32) from `normals`; returns color."""
    light = -np.asarray(light_direction, dtype=np.float32)
    light = (light / np.linalg.norm(light)).astype(np.float32)
    h, w = color.shape[:2]
    assert color.dtype == np.float32 and color.flags.c_contiguous
    lib().oracle_guro(h, w, _fp(light), _fp(color), _fp(np.ascontiguousarray(normals, dtype=np.float32)))
    return color


class OracleFiller:
    """CPU stand-in for the reference AdvancedPixelBufferFiller (pyx:20-253), n_threads=1 semantics."""

    def __init__(self, h, w, fov=90.0, z_near=0.1, z_far=1000.0, n_threads=1):
        self.h, self.w = int(h), int(w)
        self.n_threads = int(n_threads)
        self.proj_mat = projection_matrix(h, w, fov, z_near, z_far)
        self.normals_buffer = np.zeros((self.h, self.w, 3), dtype=np.float32)
        self.color_buffer = np.zeros((self.h, self.w, 3), dtype=np.float32)
        self.z_buffer = np.ones((self.h, self.w), dtype=np.float32) * np.float32(1e6)

    def get_size(self):
        return self.h, self.w

    def render_arrays(self, v, c, n):
        v = _f32_tri_array(v, "vertices")
        c = _f32_tri_array(c, "colors")
        n = _f32_tri_array(n, "normals")
        assert v.shape == c.shape == n.shape
        rc = lib().oracle_render(self.h, self.w, _fp(self.proj_mat.ravel()), _fp(v), _fp(c), _fp(n), v.shape[0],
                                 _fp(self.z_buffer), _fp(self.color_buffer), _fp(self.normals_buffer),
                                 self.n_threads)
        if rc != 0:
            raise MemoryError("oracle_render failed")

    def render_model(self, model):
        # pyx:94-96: `.copy()` on each attribute -> AttributeError when the model has no colours
        self.render_arrays(model._vertices_by_triangles.copy(), model._colors_by_triangles.copy(),
                           model._normals_by_triangles.copy())

    def get_normals_buffer(self):
        return self.normals_buffer

    def get_color_buffer(self):
        return self.color_buffer

    def get_z_buffer(self):
        return self.z_buffer
