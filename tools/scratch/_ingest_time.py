"""Timing of the model-ingest row (N4) on the GPU box: drop-in Model vs the reference's Model (oracle/_ref copy) on the
reference's assets.  Prints one JSON object.  Run: python tools/scratch/_ingest_time.py"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
OBJ = os.path.join(REF, "objects")
warnings.filterwarnings("ignore")


def timeit(fn, n):
    best = 1e9
    for _ in range(n):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best * 1e3


def main():
    import torch
    from cython3dmodelrenderer_b200 import _lib
    from cython3dmodelrenderer_b200.model import Model, parse_obj_text
    L = _lib.load_library()
    out = {}
    tex = os.path.join(OBJ, "igor_texture.png")
    for name, kw in (("T-Rex", {}), ("bunny", dict(external_texture_filename=tex))):
        path = os.path.join(OBJ, name + ".obj")
        raw = open(path, "rb").read()
        r = {"obj_bytes": len(raw)}
        Model.read_model(path, **kw)   # warm-up (CUDA context, allocator)
        r["parse_ms"] = timeit(lambda: parse_obj_text(raw), 5)
        r["read_model_ms"] = timeit(lambda: Model.read_model(path, **kw), 5)
        m = Model.read_model(path, **kw)
        r["rotate_ms"] = timeit(lambda: m.rotate([10, -80, 0]), 10)
        r["V"], r["T"] = m.n_vertices(), m.n_triangles()
        # normals kernels alone (device-resident, CUDA events)
        dv = torch.from_numpy(m._vertices).cuda()
        dt = torch.from_numpy(np.ascontiguousarray(m._triangles_vertices)).cuda()
        dn = torch.empty_like(dv)
        ws = torch.empty(L.crb_model_normals_workspace_bytes(len(dv), len(dt)), dtype=torch.uint8, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            L.crb_model_vertex_normals(dv.data_ptr(), len(dv), dt.data_ptr(), len(dt), 0, dn.data_ptr(), ws.data_ptr(), ws.numel(), s)
        e0.record()
        for _ in range(20):
            L.crb_model_vertex_normals(dv.data_ptr(), len(dv), dt.data_ptr(), len(dt), 0, dn.data_ptr(), ws.data_ptr(), ws.numel(), s)
        e1.record()
        torch.cuda.synchronize()
        r["normals_kernels_us"] = e0.elapsed_time(e1) / 20 * 1e3
        if os.path.isdir(os.path.join(REF, "crender")):
            sys.path.insert(0, REF)
            from crender.cy.data_structures import Model as Ref
            r["reference_read_model_ms"] = timeit(lambda: Ref.read_model(path, **kw), 2)
            rm = Ref.read_model(path, **kw)
            r["reference_rotate_ms"] = timeit(lambda: rm.rotate([10, -80, 0]), 2)
        out[name] = r
    # the C4 mesh family: indexed UV sphere, 10 M triangles
    from sphere_mesh import indexed_sphere
    v, tri = indexed_sphere(3200, 1564)
    dv, dt = torch.from_numpy(v).cuda(), torch.from_numpy(tri).cuda()
    dn = torch.empty_like(dv)
    ws = torch.empty(L.crb_model_normals_workspace_bytes(len(dv), len(dt)), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.crb_model_vertex_normals(dv.data_ptr(), len(dv), dt.data_ptr(), len(dt), 0, dn.data_ptr(), ws.data_ptr(), ws.numel(), s)
    e0.record()
    for _ in range(3):
        L.crb_model_vertex_normals(dv.data_ptr(), len(dv), dt.data_ptr(), len(dt), 0, dn.data_ptr(), ws.data_ptr(), ws.numel(), s)
    e1.record()
    torch.cuda.synchronize()
    out["sphere_10M"] = {"V": len(v), "T": len(tri), "normals_kernels_ms": e0.elapsed_time(e1) / 3,
                         "workspace_MB": ws.numel() / 1e6}
    out["launches"] = int(L.crb_model_launch_count())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    main()
