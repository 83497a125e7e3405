"""Generate the committed golden fixtures from the reference's own code.

Run in the build container only (needs /root/reference, via the oracle/_ref build made by
oracle/build_ref.py).  The GPU box has neither; the tests there read the files written here.

Outputs (tests/golden/):
  trex_fit.npz    -- T-Rex after the README flow (run.py:29-39): indexed arrays from which the three
                     [T,3,3] float32 inputs of render_model are rebuilt bit-exactly (v = vertices[tri_v], ...)
  bunny_fit.npz   -- bunny.obj + igor_texture.png after fit_model (SURVEY.md section 8d, config C2/C3 substitute)
  basketball_fit.npz -- basketball.obj (quads, fan-triangulated by the reference) + igor_texture.png after fit_model
  checksums.json  -- sha256 of the reference's z / colour / normal buffers (n_threads=1) for a list of
                     (model, h, w, fov) cases, plus the projection-matrix known answers
  trex_128.npz    -- full reference output buffers for T-Rex at 128x128 (small enough to commit)
"""
import hashlib
import io
import json
import os
import sys
import contextlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402

build_ref.build()
sys.path.insert(0, REF)
import numpy as np  # noqa: E402
from crender.cy.data_structures import Model  # noqa: E402
from crender.cy.pixel_buffer_filler import AdvancedPixelBufferFiller  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fit_model(m):  # run.py:30-33
    m.shift(-m.get_mean_vertex())
    m.scale(1 / m.get_max_span())
    m.shift(shift=[0, 0, 1])


def indexed(m):
    d = dict(vertices=m._vertices, normals=m._normals, tri_v=m._triangles_vertices.astype(np.int32),
             tri_n=np.asarray(m._triangles_normals, dtype=np.int32),
             colors=m._colors.astype(np.uint8), tri_vt=m._triangles_texture_coords.astype(np.int32))
    assert np.array_equal(d["colors"].astype(np.float32), m._colors)
    assert np.array_equal(d["vertices"][d["tri_v"]].view("u4"), m._vertices_by_triangles.view("u4"))
    assert np.array_equal(d["normals"][d["tri_n"]].view("u4"), m._normals_by_triangles.view("u4"))
    assert np.array_equal(d["colors"].astype(np.float32)[d["tri_vt"]], m._colors_by_triangles)
    return d


def render(m, h, w, fov, **kw):
    f = AdvancedPixelBufferFiller(h, w, fov=fov, n_threads=1, **kw)
    f.render_model(m)
    return f.get_z_buffer(), f.get_color_buffer(), f.get_normals_buffer()


def main():
    os.chdir(REF)
    trex = Model.read_model("objects/T-Rex.obj")
    trex.rotate([-90, 180, 0])
    trex.rotate([10, -80, 0])
    fit_model(trex)
    bunny = Model.read_model("objects/bunny.obj", external_texture_filename="objects/igor_texture.png")
    fit_model(bunny)
    np.savez_compressed(os.path.join(HERE, "trex_fit.npz"), **indexed(trex))
    np.savez_compressed(os.path.join(HERE, "bunny_fit.npz"), **indexed(bunny))
    # config C3 substitute (igor.obj is absent from the reference tree): the quad-faced basketball with igor's texture
    ball = Model.read_model("objects/basketball.obj", external_texture_filename="objects/igor_texture.png")
    fit_model(ball)
    np.savez_compressed(os.path.join(HERE, "basketball_fit.npz"), **indexed(ball))

    cases = {}
    models = {"trex": trex, "bunny": bunny, "basketball": ball}
    for name, h, w, fov in [("trex", 1024, 1024, 45.0), ("trex", 512, 512, 90.0), ("trex", 333, 777, 60.0),
                            ("trex", 128, 128, 45.0), ("trex", 2048, 2048, 45.0),
                            ("bunny", 1024, 1024, 45.0), ("bunny", 2048, 2048, 45.0), ("bunny", 4096, 4096, 45.0),
                            ("bunny", 500, 300, 30.0), ("basketball", 2048, 2048, 45.0), ("basketball", 1000, 1500, 70.0)]:
        m = models[name]
        z, c, n = render(m, h, w, fov)
        cases[f"{name}_{h}x{w}_fov{fov:g}"] = dict(
            model=name, h=h, w=w, fov=fov, z=sha(z), color=sha(c), normals=sha(n),
            covered=int((z < 1e5).sum()), T=int(m._vertices_by_triangles.shape[0]))
        if (name, h) == ("trex", 128):
            np.savez_compressed(os.path.join(HERE, "trex_128.npz"), z=z, color=c, normals=n)
    # compositing: bunny drawn over T-Rex in one filler (buffers persist, pyx:65-67 + no reset)
    f = AdvancedPixelBufferFiller(640, 480, fov=50.0, n_threads=1)
    f.render_model(trex)
    f.render_model(bunny)
    cases["trex_then_bunny_640x480_fov50"] = dict(
        model="trex+bunny", h=640, w=480, fov=50.0, z=sha(f.get_z_buffer()), color=sha(f.get_color_buffer()),
        normals=sha(f.get_normals_buffer()), covered=int((f.get_z_buffer() < 1e5).sum()))
    # Renderer.render with GuroIllumination (crender/cy/renderer.py:47-49, guro_illumination.py:20-27): the lit colour buffer
    from crender.cy.illumination import GuroIllumination
    for name, h, w, fov, light in [("trex", 1024, 1024, 45.0, [0, 0, 1]), ("bunny", 2048, 2048, 45.0, [0, 0, 1]),
                                   ("basketball", 777, 555, 60.0, [0.3, -0.5, 1.0]),
                                   ("bunny", 4096, 4096, 45.0, [0, 0, 1])]:       # config C2: 4096^2 "with lighting"
        z, c, n = render(models[name], h, w, fov)
        c = c.copy()
        GuroIllumination(light).draw_illumination(c, n)
        cases[f"{name}_{h}x{w}_fov{fov:g}_guro"] = dict(model=name, h=h, w=w, fov=fov, light=light, color_lit=sha(c),
                                                        covered=int((z < 1e5).sum()))
    # config C5: views of the 1024^2 T-Rex orbit; the reference is fed the camera-space arrays the view transform
    # produces (views.transform_arrays_host restates the GPU transform in NumPy float32)
    from cython3dmodelrenderer_b200 import views as VW
    for n_total, ks in [(128, (0, 32, 64, 96)), (1024, (100, 611))]:
        for k in ks:
            view = VW.orbit_views(n_total, first=k, count=1)[0]
            vk, nk = VW.transform_arrays_host(view, trex._vertices_by_triangles, trex._normals_by_triangles)
            z, c, n = render(type("M", (), dict(_vertices_by_triangles=vk, _colors_by_triangles=trex._colors_by_triangles,
                                                _normals_by_triangles=nk))(), 1024, 1024, 45.0)
            cases[f"trex_orbit{n_total}_view{k}_1024x1024_fov45"] = dict(
                model="trex", h=1024, w=1024, fov=45.0, orbit=n_total, view=k, z=sha(z), color=sha(c), normals=sha(n),
                covered=int((z < 1e5).sum()), v_in=sha(vk), n_in=sha(nk))
    # config C4 at full size: the 10 003 200-triangle UV sphere at 8192^2, rendered by the reference itself
    from cython3dmodelrenderer_b200 import synthetic
    sph = synthetic.uv_sphere(3200, 1564)
    z, c, n = render(sph, 8192, 8192, 45.0)
    cases["sphere10m_8192x8192_fov45"] = dict(
        model="uv_sphere(3200,1564)", h=8192, w=8192, fov=45.0, z=sha(z), color=sha(c), normals=sha(n),
        covered=int((z < 1e5).sum()), T=int(sph._vertices_by_triangles.shape[0]),
        v_in=sha(sph._vertices_by_triangles), c_in=sha(sph._colors_by_triangles), n_in=sha(sph._normals_by_triangles))
    del sph, z, c, n
    inputs = {k: dict(v=sha(m._vertices_by_triangles), c=sha(m._colors_by_triangles), n=sha(m._normals_by_triangles))
              for k, m in models.items()}
    with open(os.path.join(HERE, "checksums.json"), "w") as fo:
        json.dump(dict(generator="tests/golden/make_golden.py (reference Cython build, n_threads=1)",
                       numpy=np.__version__, inputs=inputs, cases=cases), fo, indent=1, sort_keys=True)
    print(json.dumps(cases, indent=1))


if __name__ == "__main__":
    with contextlib.redirect_stderr(io.StringIO()):
        main()
