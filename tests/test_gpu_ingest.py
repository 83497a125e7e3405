"""Model ingest on the GPU: the CUDA kernels behind include/crender_ingest_b200.h and the drop-in
`Model` against the ingest oracle and the golden vectors the reference's own `Model` produced.  Bit-exact (any NaN
matches any NaN, see conftest.same_f32)."""
import ctypes
import hashlib
import json
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, bits_equal, same_f32

pytestmark = pytest.mark.gpu

OBJ = os.path.join(GOLDEN, "obj")
REF_OBJ = os.path.join(ROOT, "oracle", "_ref", "objects")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def dev():
    import torch
    from cython3dmodelrenderer_b200 import _lib
    return torch, _lib.load_library()


def gpu_normals(dev, v, tri, invert=False):
    torch, L = dev
    from cython3dmodelrenderer_b200._lib import check
    from oracle import ingest as I
    v = np.ascontiguousarray(v, np.float32)
    tri = I.wrap_indices(tri, len(v))
    dv, dt = torch.from_numpy(v).cuda(), torch.from_numpy(tri).cuda()
    out = torch.full((len(v), 3), 7.0, dtype=torch.float32, device="cuda")
    ws = torch.empty(L.crb_model_normals_workspace_bytes(len(v), len(tri)), dtype=torch.uint8, device="cuda")
    before = L.crb_model_launch_count()
    check(L.crb_model_vertex_normals(dv.data_ptr(), len(v), dt.data_ptr(), len(tri), int(invert), out.data_ptr(),
                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert L.crb_model_launch_count() - before >= (6 if len(tri) else 4)
    return out.cpu().numpy()


def golden(name):
    d = np.load(os.path.join(GOLDEN, f"ingest_{name}.npz"))
    return d, json.loads(str(d["hashes"]))


@pytest.mark.parametrize("name", ["quirks", "torus", "torus_inv", "fan", "cube_pm"])
def test_vertex_normals_equal_reference_golden(dev, name):
    g, _ = golden(name)
    for stage in ("read", "rot"):
        n = gpu_normals(dev, g[f"{stage}_vertices"], g[f"{stage}_triangles_vertices"],
                        invert=(name == "torus_inv" and stage == "read"))
        assert same_f32(n, g[f"{stage}_normals"]), f"{name}/{stage}"


@pytest.mark.parametrize("seed", range(12))
def test_vertex_normals_equal_oracle_on_random_meshes(dev, seed):
    from oracle import ingest as I
    rng = np.random.default_rng(100 + seed)
    V = int(rng.integers(1, 3000))
    T = int(rng.integers(0, 4 * V + 2))
    v = rng.standard_normal((V, 3)).astype(np.float32)
    if seed % 3 == 0:
        v = (np.round(v * 2) / 2).astype(np.float32)     # many exactly repeated normals, degenerate faces, -0.0
    if seed % 4 == 1:
        v[rng.integers(0, V, 3)] = np.float32(np.inf)
    tri = rng.integers(-V, V, (T, 3)).astype(np.int32)
    if seed % 5 == 2 and T > 10:
        tri[: T // 2, 0] = 0                              # a vertex of very high valence
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = I.vertex_normals(v, tri, invert=bool(seed % 2))
    got = gpu_normals(dev, v, tri, invert=bool(seed % 2))
    assert same_f32(got, want)


def test_vertex_normals_long_lists_and_determinism(dev):
    """5 000 distinct face normals around one vertex (kept list far beyond the 32 a warp holds in registers) plus exact
    repeats; scattered incidence order must not matter: repeated runs are bit-identical."""
    from oracle import ingest as I
    rng = np.random.default_rng(5)
    n = 5000
    a = np.sort(rng.uniform(0, 2 * np.pi, n + 1))
    ring = np.stack([np.cos(a), np.sin(a), 2 + 0.2 * rng.standard_normal(n + 1)], 1)
    v = np.concatenate([[[0, 0, 2.5]], ring]).astype(np.float32)
    tri = np.stack([np.zeros(n, int), np.arange(1, n + 1), np.arange(2, n + 2)], 1)
    tri = np.concatenate([tri, tri[::7]]).astype(np.int32)
    want = I.vertex_normals(v, tri)
    runs = [gpu_normals(dev, v, tri) for _ in range(3)]
    assert bits_equal(runs[0], want)
    assert bits_equal(runs[1], runs[0]) and bits_equal(runs[2], runs[0])


def test_vertex_normals_kept_list_beyond_shared_memory(dev):
    """14 000 distinct normals around one vertex: the heavy-vertex kernel's kept list spills from shared memory
    (12 288 normals) into its global scratch; a second heavy vertex shares the launch."""
    from oracle import ingest as I
    rng = np.random.default_rng(6)
    n = 14000
    a = np.sort(rng.uniform(0, 2 * np.pi, n + 1))
    ring = np.stack([np.cos(a), np.sin(a), 2 + 0.2 * rng.standard_normal(n + 1)], 1)
    v = np.concatenate([[[0, 0, 2.5]], ring, [[0, 0, 1.0]]]).astype(np.float32)
    tri = np.stack([np.zeros(n, int), np.arange(1, n + 1), np.arange(2, n + 2)], 1)
    low = np.stack([np.full(900, n + 2), np.arange(1, 901), np.arange(2, 902)], 1)
    tri = np.concatenate([tri, low, tri[::3]]).astype(np.int32)
    assert bits_equal(gpu_normals(dev, v, tri), I.vertex_normals(v, tri))


def test_vertex_normals_sphere_scale(dev):
    """A 2 M-triangle UV sphere (the C4 mesh family at 1/5 scale): indexed, poles of valence 1 600."""
    from oracle import ingest as I
    from sphere_mesh import indexed_sphere
    v, tri = indexed_sphere(1600, 626)
    assert bits_equal(gpu_normals(dev, v, tri), I.vertex_normals(v, tri))


def test_vertex_colors_and_gather_equal_oracle(dev):
    torch, L = dev
    from cython3dmodelrenderer_b200._lib import check
    from oracle import ingest as I
    rng = np.random.default_rng(9)
    for width, (h, w) in ((2, (48, 64)), (3, (1, 1)), (2, (513, 7))):
        n = 4000
        vt = rng.uniform(-0.5, 1.5, (n, width)).astype(np.float32)
        vt[:8, 0] = [np.nan, np.inf, -np.inf, 3e9, -3e9, 1.0, 0.0, -0.0]
        vt[8:16, 1] = [np.nan, np.inf, -np.inf, 3e9, -3e9, 1.0, 0.0, -0.0]
        tex = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out = torch.empty((n, 3), dtype=torch.float32, device="cuda")
        d_vt, d_tex = torch.from_numpy(vt).cuda(), torch.from_numpy(tex).cuda()   # kept alive across the launch
        check(L.crb_model_vertex_colors(d_vt.data_ptr(), n, width, d_tex.data_ptr(), h, w, out.data_ptr(), None))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = I.vertex_colors(vt, tex)
        assert bits_equal(out.cpu().numpy(), want)
        tri = rng.integers(0, n, (5000, 3)).astype(np.int32)
        by = torch.empty((5000, 3, 3), dtype=torch.float32, device="cuda")
        d_tri = torch.from_numpy(tri).cuda()
        check(L.crb_model_gather(out.data_ptr(), d_tri.data_ptr(), 5000, by.data_ptr(), None))
        assert bits_equal(by.cpu().numpy(), want[tri])


def test_bad_arguments_fail_loudly(dev):
    torch, L = dev
    from cython3dmodelrenderer_b200 import _lib
    v = torch.zeros((4, 3), device="cuda")
    t = torch.zeros((2, 3), dtype=torch.int32, device="cuda")
    assert L.crb_model_vertex_normals(v.data_ptr(), 4, t.data_ptr(), 2, 0, v.data_ptr(), None, 0, None) == _lib.CRB_ERR_STATE
    assert b"workspace" in L.crb_last_error()
    assert L.crb_model_vertex_colors(v.data_ptr(), 4, 1, v.data_ptr(), 4, 4, v.data_ptr(), None) == _lib.CRB_ERR_INVALID
    assert L.crb_model_gather(None, t.data_ptr(), 2, v.data_ptr(), None) == _lib.CRB_ERR_INVALID


# ------------------------------------------------------------------------------------------------- drop-in Model
def check_stage(m, g, hashes, stage):
    for f in ("_vertices", "_normals", "_colors", "_texture_coords", "_mean_vertex"):
        if f"{stage}{f}" in g:
            assert same_f32(getattr(m, f), g[f"{stage}{f}"]), f"{stage}{f}"
    for f in ("_triangles_vertices", "_triangles_normals", "_triangles_texture_coords"):
        if f"{stage}{f}" in g:
            assert np.array_equal(getattr(m, f), g[f"{stage}{f}"]), f"{stage}{f}"
    assert same_f32(np.float32(m._max_span), g[f"{stage}_max_span"])
    nan = np.isnan(m._normals).any() or np.isnan(m._vertices).any()
    for f in ("_vertices_by_triangles", "_normals_by_triangles", "_colors_by_triangles"):
        a = getattr(m, f)
        if f"{stage}{f}" in hashes:
            assert a.dtype == np.float32 and a.shape == (m.n_triangles(), 3, 3)
            if not nan:
                assert sha(a) == hashes[f"{stage}{f}"], f"{stage}{f}"
        else:
            assert a is None


@pytest.mark.parametrize("name", ["quirks", "torus", "torus_inv", "torus_ext", "fan", "cube_pm"])
def test_model_reproduces_reference_model(name):
    """read_model -> rotate -> the README's fit_model, every attribute of every stage against the reference's."""
    from cython3dmodelrenderer_b200.model import Model
    g, hashes = golden(name)
    kw = {}
    if name == "torus_inv":
        kw["invert_calculated_normals"] = True
    if name == "torus_ext":
        kw["external_texture_filename"] = os.path.join(OBJ, "checker.png")
    base = name.split("_")[0] if name.startswith("torus") else name
    cwd = os.getcwd()
    os.chdir(OBJ)   # mtllib paths are relative to the directory part of the name given, as upstream
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = Model.read_model(base + ".obj", **kw)
    finally:
        os.chdir(cwd)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        check_stage(m, g, hashes, "read")
        m.rotate([10, -80, 0])
        check_stage(m, g, hashes, "rot")
        m.shift(-m.get_mean_vertex())
        m.scale(1 / m.get_max_span())
        m.shift(shift=[0, 0, 1])
        check_stage(m, g, hashes, "fit")
    dv, dc, dn = m.device_triangles()
    assert bits_equal(dv.cpu().numpy(), m._vertices_by_triangles) or np.isnan(m._vertices).any()
    assert (dc is None) == (m._colors_by_triangles is None)


def test_model_on_the_reference_assets_and_render(trex):
    """T-Rex / bunny / basketball from the reference's objects directory (shipped with oracle/_ref, not committed):
    checksums of every stage; then the README flow end to end -- drop-in Model -> drop-in filler -> the golden frame."""
    if not os.path.isdir(REF_OBJ):
        pytest.skip("oracle/_ref/objects not present")
    from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller
    from cython3dmodelrenderer_b200.model import Model
    sums = json.load(open(os.path.join(GOLDEN, "ingest_checksums.json")))
    tex = os.path.join(REF_OBJ, "igor_texture.png")
    for name, kw in (("T-Rex", {}), ("basketball", dict(external_texture_filename=tex)),
                     ("bunny", dict(external_texture_filename=tex))):
        m = Model.read_model(os.path.join(REF_OBJ, name + ".obj"), **kw)
        for stage in ("read", "rot", "fit"):
            if stage == "rot":
                m.rotate([10, -80, 0])
            if stage == "fit":
                m.shift(-m.get_mean_vertex())
                m.scale(1 / m.get_max_span())
                m.shift(shift=[0, 0, 1])
            for f in ("_vertices", "_normals", "_colors", "_vertices_by_triangles", "_normals_by_triangles",
                      "_colors_by_triangles"):
                assert sha(getattr(m, f)) == sums[name][f"{stage}{f}"]["sha256_16"], f"{name} {stage}{f}"
    # README flow (run.py:30-39) on the drop-in classes only
    m = Model.read_model(os.path.join(REF_OBJ, "T-Rex.obj"))
    m.rotate([-90, 180, 0])
    m.rotate([10, -80, 0])
    m.shift(-m.get_mean_vertex())
    m.scale(1 / m.get_max_span())
    m.shift(shift=[0, 0, 1])
    assert bits_equal(m._vertices_by_triangles, trex._vertices_by_triangles)
    assert bits_equal(m._normals_by_triangles, trex._normals_by_triangles)
    assert bits_equal(m._colors_by_triangles, trex._colors_by_triangles)
    f = AdvancedPixelBufferFiller(1024, 1024, fov=45)
    f.render_model(m)
    chk = json.load(open(os.path.join(GOLDEN, "checksums.json")))
    z = f.get_z_buffer()
    assert int((z < 1e5).sum()) == 252539
    assert hashlib.sha256(z.tobytes()).hexdigest() == chk["cases"]["trex_1024x1024_fov45"]["z"]   # the reference's own frame
    g = AdvancedPixelBufferFiller(1024, 1024, fov=45)
    g.render_arrays(*m.device_triangles())           # device-resident twins: no host round trip
    assert bits_equal(g.get_z_buffer(), z) and bits_equal(g.get_color_buffer(), f.get_color_buffer())
