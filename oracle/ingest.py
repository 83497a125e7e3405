"""CPU oracle of the model-ingest row: ctypes front-end of oracle/ingest_oracle.c plus a
line-by-line Python restatement of the reference's OBJ reader.

TEST INFRASTRUCTURE ONLY: imported by tests/ and __graft_entry__.smoke().  The product package never imports it.
Parity status: pinned against the reference's own `Model` (crender/cy/data_structures/model.py) -- see
tests/test_ingest_oracle.py and tests/golden/make_golden_ingest.py.  `model.py` below is that file.
"""
import ctypes

import numpy as np

from . import oracle as _O

_ready = False


def _lib():
    global _ready
    L = _O.lib()
    if not _ready:
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        L.ingest_face_normal.argtypes = [fp, fp, fp, fp]
        L.ingest_face_normal.restype = None
        L.ingest_vertex_normals.argtypes = [fp, ctypes.c_int64, ip, ctypes.c_int64, ctypes.c_int, fp]
        L.ingest_vertex_normals.restype = ctypes.c_int
        L.ingest_vertex_colors.argtypes = [fp, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_uint8),
                                           ctypes.c_int, ctypes.c_int, fp]
        L.ingest_vertex_colors.restype = None
        L.ingest_gather.argtypes = [fp, ip, ctypes.c_int64, fp]
        L.ingest_gather.restype = None
        _ready = True
    return L


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def wrap_indices(tri, n):
    """NumPy fancy indexing / Python list indexing of negative indices (model.py:158,172,179-181)."""
    tri = np.asarray(tri, dtype=np.int64)
    if tri.size and (tri.min() < -n or tri.max() >= n):
        raise IndexError(f"index out of bounds for axis 0 with size {n}")
    return np.ascontiguousarray(np.where(tri < 0, tri + n, tri).astype(np.int32))


def face_normal(tri):
    """model.py:196-201 on one [3,3] float32 triangle."""
    t = np.ascontiguousarray(tri, dtype=np.float32)
    n = np.zeros(3, np.float32)
    _lib().ingest_face_normal(_fp(t[0]), _fp(t[1]), _fp(t[2]), _fp(n))
    return n


def vertex_normals(vertices, tri, invert=False):
    """model.py:174-188 (+168-169): [V,3] float32, [T,3] int -> [V,3] float32."""
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    t = wrap_indices(tri, len(v))
    out = np.zeros((len(v), 3), np.float32)
    if _lib().ingest_vertex_normals(_fp(v), len(v), _ip(t), len(t), int(bool(invert)), _fp(out)) != 0:
        raise MemoryError
    return out


def vertex_colors(texture_coords, texture):
    """model.py:147-150: per-vertex nearest texel, float32 [n,3] (BGR)."""
    vt = np.ascontiguousarray(texture_coords, dtype=np.float32)
    tex = np.ascontiguousarray(texture, dtype=np.uint8)
    h, w, _ = tex.shape
    out = np.zeros((len(vt), 3), np.float32)
    _lib().ingest_vertex_colors(_fp(vt), len(vt), vt.shape[1], tex.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                h, w, _fp(out))
    return out


def gather(attr, tri):
    """model.py:151,158,172: attr[tri] -> [T,3,3]."""
    a = np.ascontiguousarray(attr, dtype=np.float32)
    t = wrap_indices(tri, len(a))
    out = np.zeros((len(t), 3, 3), np.float32)
    _lib().ingest_gather(_fp(a), _ip(t), len(t), _fp(out))
    return out


# ---------------------------------------------------------------------------------------------------------------
# OBJ reader, restated line by line (model.py:7-77 read_model, 258-312 the _read_* helpers).  Small inputs only.

def _fix(index):
    return index - 1 if index > 0 else index   # model.py:275-279: 1-based -> 0-based, 0 and negatives kept


def parse_obj(text):
    """Returns dict(vertices, texture_coords, normals, tri_v, tri_vt, tri_vn, mtllibs); tri_vt / tri_vn are None once
    any face lacked them (model.py:48-56).  Lines that raise are skipped, as with silent=True (model.py:71-74).
    `text` is the file content as Python's text mode would hand it over (universal newlines already applied)."""
    vertices, texture_coords, normals = [], [], []
    tri_v, tri_vt, tri_vn, mtllibs = [], [], [], []
    # text-mode iteration: universal newlines, then lines end at '\n' only (not at \v, \f, ... like str.splitlines)
    text = text.replace('\r\n', '\n').replace('\r', '\n')
    for line in text.split('\n'):
        line += '\n'   # a missing final newline only matters for an empty `mtllib ` payload
        try:
            if line == '' or line[0] == '#':
                continue
            parts = line.split(' ', 1)
            if len(parts) != 2:
                continue
            command, data = parts
            if command == 'v':
                c = [float(t) for t in data.split()]
                assert len(c) >= 3
                vertices.append(c[:3])
            elif command == 'vt':
                texture_coords.append([float(t) for t in data.split()])
            elif command == 'vn':
                c = [float(t) for t in data.split()]
                assert len(c) == 3
                normals.append(c)
            elif command == 'f':
                comp = data.split()
                vs, vts, vns = [], [], []
                for i in range(len(comp) - 2):
                    tv, tvt, tvn = [], [], []
                    for corner in (comp[0], comp[1 + i], comp[2 + i]):
                        a, b, c = (corner + '//').split('/')[:3]
                        tv.append(_fix(int(a)))
                        if b == '':
                            tvt = None
                        if tvt is not None:
                            tvt.append(_fix(int(b)))
                        if c == '':
                            tvn = None
                        if tvn is not None:
                            tvn.append(_fix(int(c)))
                    vs.append(tv)
                    vts.append(tvt)
                    vns.append(tvn)
                tri_v.extend(vs)
                if vts.count(None) > 0:
                    tri_vt = None
                if tri_vt is not None:
                    tri_vt.extend(vts)
                if vns.count(None) > 0:
                    tri_vn = None
                if tri_vn is not None:
                    tri_vn.extend(vns)
            elif command == 'mtllib':
                mtllibs.append(data)
        except Exception:
            pass
    return dict(vertices=vertices, texture_coords=texture_coords, normals=normals,
                tri_v=tri_v, tri_vt=tri_vt, tri_vn=tri_vn, mtllibs=mtllibs)
