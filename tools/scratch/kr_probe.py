"""k_raster probe for kernel experiments (GPU box): for the headline batch (T-Rex 1024^2, 128 orbit views per launch), the
C2 frame (bunny 4096^2 + fused Guro) and the C4 frame (10 M-triangle sphere, 8192^2) prints the mean k_raster launch
(CUDA events on the launching stream), its algorithmic GB/s and fraction of the measured HBM peak, and the whole step.
usage: kr_probe.py [label] [workloads: t,b,s]   (CRB_LIB_OVERRIDE selects another build of the library)"""
import json, os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import numpy as np
import torch
from conftest import load_indexed
from cython3dmodelrenderer_b200 import AdvancedPixelBufferFiller, views as VW, synthetic, _lib
label = sys.argv[1] if len(sys.argv) > 1 else "-"
which = sys.argv[2] if len(sys.argv) > 2 else "tbs"
PEAK = 6547.2
try:
    PEAK = float(json.load(open(os.path.join(_ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {"label": label}


def timed(f, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    f.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    f.join()
    e1.record()
    torch.cuda.synchronize()
    n, ms = f.profile_read()
    f.profile(False)
    return e0.elapsed_time(e1) / reps, ms / max(n, 1), n / reps


if "t" in which:
    m = load_indexed("trex"); res, V = 1024, 128
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
    views = torch.from_numpy(VW.orbit_views(V)).cuda()
    z = torch.empty((V, res, res), device="cuda"); c = torch.empty((V, res, res, 3), device="cuda"); n = torch.empty((V, res, res, 3), device="cuda")
    f = AdvancedPixelBufferFiller(res, res, fov=45.0)
    f.render_views(dv, dc, dn, views, z_out=z, color_out=c, normals_out=n, chunk=V)
    step, k, _ = timed(f, lambda: f.render_views(dv, dc, dn, views, z_out=z, color_out=c, normals_out=n, chunk=V, check_status=False, defer_join=True), 20)
    B = (108 * dv.shape[0] + 28 * res * res) * V
    out["trex128"] = {"step_ms": round(step, 4), "k_ms": round(k, 4), "frac": round(B / k / 1e6 / PEAK, 4), "fps": round(V / step * 1000), "covered0": int((z[0] < 1e5).sum())}
    _lib.check(f._L.crb_set_option(f._handle, _lib.CRB_OPT_CHUNK_PIPELINE, 0))
    step, k, _ = timed(f, lambda: f.render_views(dv, dc, dn, views, z_out=z, color_out=c, normals_out=n, chunk=V, check_status=False), 10)
    out["trex128"]["k_ms_alone"] = round(k, 4); out["trex128"]["frac_alone"] = round(B / k / 1e6 / PEAK, 4); out["trex128"]["step_ms_serial"] = round(step, 4)
    # one frame at a time (C1 latency)
    g = AdvancedPixelBufferFiller(res, res, fov=45.0)
    g.clear(); g.render_arrays(dv, dc, dn)

    def one():
        g._pending_clear = True
        g.render_arrays(dv, dc, dn, check_status=False)
    step, k, _ = timed(g, one, 200)
    out["trex1"] = {"step_us": round(step * 1000, 2), "k_us": round(k * 1000, 2)}
    del f, g, z, c, n
if "b" in which:
    m = load_indexed("bunny"); res = 4096
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
    f = AdvancedPixelBufferFiller(res, res, fov=45.0)
    zb, cb, nb = f.device_buffers()
    ident = torch.from_numpy(VW.view_matrix()[None, :]).cuda()
    f.render_views(dv, dc, dn, ident, z_out=zb[None], color_out=cb[None], normals_out=nb[None], guro_light=[0, 0, 1], chunk=1)
    step, k, _ = timed(f, lambda: f.render_views(dv, dc, dn, ident, z_out=zb[None], color_out=cb[None], normals_out=nb[None], guro_light=[0, 0, 1], chunk=1, check_status=False), 50)
    B = 108 * dv.shape[0] + 28 * res * res
    out["bunny4096guro"] = {"step_ms": round(step, 4), "k_ms": round(k, 4), "frac": round(B / k / 1e6 / PEAK, 4), "covered": int((zb < 1e5).sum())}
    del f
if "s" in which:
    m = synthetic.uv_sphere(3200, 1564); res = 8192
    dv, dc, dn = (torch.from_numpy(a).cuda() for a in (m._vertices_by_triangles, m._colors_by_triangles, m._normals_by_triangles))
    f = AdvancedPixelBufferFiller(res, res, fov=45.0)
    f.clear(); f.render_arrays(dv, dc, dn)

    def one():
        f._pending_clear = True
        f.render_arrays(dv, dc, dn, check_status=False)
    step, k, _ = timed(f, one, 20)
    B = 108 * dv.shape[0] + 28 * res * res
    out["sphere8192"] = {"step_ms": round(step, 4), "k_ms": round(k, 4), "frac": round(B / k / 1e6 / PEAK, 4), "covered": int((f.device_buffers()[0] < 1e5).sum())}
print(json.dumps(out))
