import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the reference's own Cython build in oracle/_ref")


class TriModel:
    """Duck-typed stand-in for crender.cy.data_structures.Model: render_model only reads these three attributes."""

    def __init__(self, v, c, n):
        self._vertices_by_triangles = v
        self._colors_by_triangles = c
        self._normals_by_triangles = n


def load_indexed(name):
    d = np.load(os.path.join(GOLDEN, name + "_fit.npz"))
    v = np.ascontiguousarray(d["vertices"][d["tri_v"]])
    n = np.ascontiguousarray(d["normals"][d["tri_n"]])
    c = np.ascontiguousarray(d["colors"].astype(np.float32)[d["tri_vt"]])
    return TriModel(v, c, n)


def case_model(info):
    """The model a golden case of checksums.json was rendered from (tests/golden/make_golden.py): a fixture as it is, a
    view of the T-Rex orbit (camera-space arrays of views.transform_arrays_host, config C5), or the full-size UV sphere
    of config C4.  Returns (model, n_threads for the CPU oracle)."""
    if "orbit" in info:
        from cython3dmodelrenderer_b200 import views as VW
        base = load_indexed("trex")
        view = VW.orbit_views(info["orbit"], first=info["view"], count=1)[0]
        vk, nk = VW.transform_arrays_host(view, base._vertices_by_triangles, base._normals_by_triangles)
        return TriModel(vk, base._colors_by_triangles, nk), 1
    if info["model"].startswith("uv_sphere"):
        from cython3dmodelrenderer_b200 import synthetic
        m = synthetic.uv_sphere(3200, 1564)
        return TriModel(m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles), os.cpu_count() or 1
    return load_indexed(info["model"]), 1


def random_scene(seed, T=None, span=1.0):
    """SURVEY 8d property-test scenes: mixed sizes, depths 0.2..5, some behind the camera, some snapped to a grid."""
    rng = np.random.default_rng(seed)
    T = int(rng.integers(1, 400)) if T is None else T
    ctr = rng.uniform(-span, span, (T, 1, 3)).astype(np.float32)
    ctr[..., 2] = rng.uniform(0.2, 5, (T, 1))
    ext = np.exp(rng.uniform(np.log(0.002), np.log(1.0), (T, 1, 1))).astype(np.float32)
    v = (ctr + rng.uniform(-1, 1, (T, 3, 3)).astype(np.float32) * ext).astype(np.float32)
    if seed % 5 == 0:
        v[..., 2] -= np.float32(1.0)
    if seed % 7 == 0:
        v = (np.round(v * 8) / 8).astype(np.float32)
    v[v[..., 2] == 0] += np.float32(0.01)
    n = rng.standard_normal((T, 3, 3)).astype(np.float32)
    c = (rng.random((T, 3, 3)) * 255).astype(np.float32)
    return TriModel(v, c, n)


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def same_f32(a, b):
    """Bit-equal except that any NaN matches any NaN: NaN sign / payload bits depend on the instruction set (x86 SSE makes
    0xFFC00000, the GPU 0x7FFFFFFF) and mean nothing to NumPy code.  Used by the model-ingest tests."""
    a, b = np.ascontiguousarray(a, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


@pytest.fixture(scope="session")
def trex():
    return load_indexed("trex")


@pytest.fixture(scope="session")
def bunny():
    return load_indexed("bunny")


@pytest.fixture(scope="session")
def basketball():
    return load_indexed("basketball")
