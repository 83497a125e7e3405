"""Golden vectors of the model-ingest row, made by the REFERENCE's own `Model`
(crender/cy/data_structures/model.py, imported from oracle/_ref = a build-only copy of /root/reference).

Run in the build container:  python tests/golden/make_golden_ingest.py
Writes tests/golden/ingest_<name>.npz for the committed fixtures of tests/golden/obj (full arrays) and
tests/golden/ingest_checksums.json for the reference's own assets (sha256 of the arrays; the .obj files themselves are
not committed -- tests read them from oracle/_ref/objects when that is present).

Per model: the state after read_model, after rotate([10,-80,0]) (normals recomputed) and after the README's fit_model
(shift / scale / shift).  NumPy 2.3.5 + OpenBLAS 0.3.30 of this image.
"""
import hashlib
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
warnings.filterwarnings("ignore")
from crender.cy.data_structures import Model  # noqa: E402

FIELDS = ["_vertices", "_normals", "_colors", "_texture_coords", "_triangles_vertices", "_triangles_normals",
          "_triangles_texture_coords", "_vertices_by_triangles", "_normals_by_triangles", "_colors_by_triangles",
          "_mean_vertex", "_max_span"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def snapshot(m, tag, out):
    for f in FIELDS:
        a = getattr(m, f)
        if a is not None:
            out[f"{tag}{f}"] = np.asarray(a)


def stages(path, **kw):
    out = {}
    m = Model.read_model(path, **kw)
    snapshot(m, "read", out)
    m.rotate([10, -80, 0])
    snapshot(m, "rot", out)
    m.shift(-m.get_mean_vertex())
    m.scale(1 / m.get_max_span())
    m.shift(shift=[0, 0, 1])
    snapshot(m, "fit", out)
    return out


def main():
    obj = os.path.join(HERE, "obj")
    cases = {
        "quirks": dict(path=os.path.join(obj, "quirks.obj")),
        "torus": dict(path=os.path.join(obj, "torus.obj")),
        "torus_inv": dict(path=os.path.join(obj, "torus.obj"), invert_calculated_normals=True),
        "torus_ext": dict(path=os.path.join(obj, "torus.obj"),
                          external_texture_filename=os.path.join(obj, "checker.png")),
        "fan": dict(path=os.path.join(obj, "fan.obj")),
        "cube_pm": dict(path=os.path.join(obj, "cube_pm.obj")),
    }
    for name, kw in cases.items():
        path = kw.pop("path")
        cwd = os.getcwd()
        os.chdir(os.path.dirname(path))   # mtllib paths are relative to the .obj's directory string
        try:
            out = stages(os.path.basename(path) if name != "torus_ext" else path, **kw)
        finally:
            os.chdir(cwd)
        # by-triangle arrays are re-derivable (attr[tri]); keep them out of the files, keep their hashes
        small = {k: v for k, v in out.items() if "_by_triangles" not in k}
        small["hashes"] = np.array(json.dumps({k: sha(v) for k, v in out.items()}))
        np.savez_compressed(os.path.join(HERE, f"ingest_{name}.npz"), **small)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("read")})
    sums = {}
    ref_obj = os.path.join(ROOT, "oracle", "_ref", "objects")
    for name, kw in {"T-Rex": {}, "basketball": dict(external_texture_filename=os.path.join(ref_obj, "igor_texture.png")),
                     "bunny": dict(external_texture_filename=os.path.join(ref_obj, "igor_texture.png")),
                     "cube": {}, "Cube2": {}}.items():
        out = stages(os.path.join(ref_obj, name + ".obj"), **kw)
        sums[name] = {k: {"shape": list(v.shape), "dtype": str(v.dtype), "sha256_16": sha(v)} for k, v in out.items()}
        print(name, len(out))
    json.dump(sums, open(os.path.join(HERE, "ingest_checksums.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
