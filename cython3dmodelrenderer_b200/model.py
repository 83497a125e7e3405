"""Drop-in `Model`: the reference's mesh container
(crender/cy/data_structures/model.py -- "model.py" below) with its slow parts moved behind the C ABI
(include/crender_ingest_b200.h):

* `read_model` -- the .obj text is read by `crb_obj_parse` (host C++, model.py:7-77 and 258-312 restated);
* smooth vertex normals (`_compute_normals_by_vertex`, model.py:174-188: a Python loop over triangles upstream, 0.7 s
  per call on T-Rex and run again by every `rotate`) -- `crb_model_vertex_normals`, CUDA;
* per-vertex texture colours (model.py:147-150) -- `crb_model_vertex_colors`, CUDA;
* the `attr[triangles]` gathers that build `_vertices_by_triangles / _colors_by_triangles / _normals_by_triangles`
  (model.py:151,158,172) -- `crb_model_gather`, CUDA; the results also stay on the device (`device_triangles()`), so a
  filler can render them without another host round trip.

Same constructor, methods, attribute names and NumPy results (bit for bit) as upstream.  The O(V) vectorised NumPy
expressions of upstream (`mean`, `max_span`, the 3x3 rotation product) are kept as NumPy on the host: they are not
the cost.  PyTorch only owns device memory.  There is no CPU fallback for the CUDA parts.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("cython3dmodelrenderer_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def parse_obj_text(data: bytes):
    """`crb_obj_parse` on the bytes of an .obj file -> dict of NumPy arrays (host only, no GPU needed).

    vertices [n,3] f32, texture_coords [n,width] f32, normals [n,3] f32, tri_v / tri_vt / tri_vn [T,3] int32
    (tri_vt / tri_vn None once a face lacked them, model.py:48-56), mtllibs (list of str), bad_lines (lines Python
    would have raised on), first_bad_line (upstream's `line_index + 1` of the first of them, or 0)."""
    L = _lib.load_library()
    handle = ctypes.c_void_p()
    check(L.crb_obj_parse(data, len(data), ctypes.byref(handle)))
    try:
        counts = (ctypes.c_int64 * 10)()
        check(L.crb_obj_counts(handle, counts))
        n_v, n_vt, width, n_vn, n_tri, has_vt, has_vn, n_mtl, n_bad, first_bad = (int(c) for c in counts)
        if width < 0:
            # np.array(texture_coords, dtype=np.float32) on rows of different lengths (model.py:143)
            raise ValueError("setting an array element with a sequence. The requested array has an inhomogeneous "
                             "shape after 1 dimensions.")
        out = {
            "vertices": np.empty((n_v, 3), np.float32),
            "texture_coords": np.empty((n_vt, max(width, 0)), np.float32),
            "normals": np.empty((n_vn, 3), np.float32),
            "tri_v": np.empty((n_tri, 3), np.int32),
            "tri_vt": np.empty((n_tri, 3), np.int32) if has_vt else None,
            "tri_vn": np.empty((n_tri, 3), np.int32) if has_vn else None,
        }
        ptr = lambda a: None if a is None or a.size == 0 else a.ctypes.data   # noqa: E731
        check(L.crb_obj_copy(handle, ptr(out["vertices"]), ptr(out["texture_coords"]), ptr(out["normals"]),
                             ptr(out["tri_v"]), ptr(out["tri_vt"]), ptr(out["tri_vn"])))
        mtllibs = []
        for k in range(n_mtl):
            p, n = ctypes.c_char_p(), ctypes.c_size_t()
            check(L.crb_obj_mtllib(handle, k, ctypes.byref(p), ctypes.byref(n)))
            mtllibs.append(ctypes.string_at(p, n.value).decode("utf-8", "replace"))
        out["mtllibs"] = mtllibs
        out["bad_lines"] = n_bad
        out["first_bad_line"] = first_bad
        return out
    finally:
        L.crb_obj_free(handle)


def _wrap_indices(tri, n):
    """What NumPy fancy indexing (and the Python list indexing of model.py:179-181) does with negative indices."""
    tri = np.asarray(tri)
    if tri.size and (int(tri.min()) < -n or int(tri.max()) >= n):
        bad = int(tri.max()) if int(tri.max()) >= n else int(tri.min())
        raise IndexError(f"index {bad} is out of bounds for axis 0 with size {n}")
    return np.ascontiguousarray(np.where(tri < 0, tri + n, tri), dtype=np.int32)


class Model:
    """model.py:5.  Extra keyword (not upstream): `device` -- CUDA device index (default: torch's current device)."""

    # ------------------------------------------------------------------------------------------------ reading
    @staticmethod
    def read_model(filename: str, silent=True, external_texture_filename=None,
                   recalculate_normals=True, invert_calculated_normals=False, device=None):
        """model.py:7-77."""
        texture = Model._read_texture_file(external_texture_filename) if external_texture_filename is not None else None
        with open(filename.strip(), 'rb') as f:
            raw = f.read()
        if not raw.isascii():
            raw.decode('utf-8')   # upstream reads in text mode: undecodable bytes raise UnicodeDecodeError there too
        parsed = parse_obj_text(raw)
        if not silent and parsed["bad_lines"]:
            raise RuntimeError(f'Error occurred while parsing line #{parsed["first_bad_line"]} of "{filename}"')
        for data in parsed["mtllibs"]:
            # model.py:58-65 -- only while no texture is known; failures are swallowed like any other line's
            if texture is not None:
                continue
            try:
                data = data + '\n'
                path = (Model._get_dir(filename) if data[0] != '/' else '') + data
                image_filename = Model._read_material_file(path, filename.strip())
                texture = None
                if image_filename is not None:
                    image_filename = (Model._get_dir(filename) if image_filename[0] != '/' else '') + image_filename
                    texture = Model._read_texture_file(image_filename)
            except Exception as e:
                if not silent:
                    raise RuntimeError(f'Error occurred while parsing a mtllib line of "{filename}"') from e
        return Model(parsed["vertices"], parsed["tri_v"],
                     parsed["texture_coords"], parsed["tri_vt"], texture,
                     parsed["normals"], parsed["tri_vn"], recalculate_normals, invert_calculated_normals,
                     device=device)

    @staticmethod
    def _read_material_file(filename, origin):
        """model.py:80-112: the last `map_Kd` payload of the .mtl file, or None."""
        image_filename = None
        try:
            with open(filename.strip(), 'r') as f:
                for line in f:
                    if line == '' or line[0] == '#':
                        continue
                    parts = line.split(' ', 1)
                    if len(parts) == 2 and parts[0] == 'map_Kd':
                        image_filename = parts[1]
        except Exception as e:
            print(f"Error occurred while parsing material file of object file '{origin}':")
            print(e)
            print('Material info will be ignored')
        return image_filename

    @staticmethod
    def _read_texture_file(filename):
        import cv2   # upstream's image reader (model.py:115); BGR uint8 [h,w,3] or None
        return cv2.imread(filename.strip())

    @staticmethod
    def _get_dir(filename):
        head, sep, _ = filename.rpartition('/')
        return head + '/' if sep else ''

    # ------------------------------------------------------------------------------------------------ building
    def __init__(self, vertices, triangles_vertices,
                 texture_coords=None, triangles_texture_coords=None, texture=None,
                 normals=None, triangles_normals=None, recalculate_normals=True, invert_calculated_normals=False,
                 device=None):
        """model.py:117-151."""
        torch = _torch()
        self._L = _lib.load_library()
        self._torch = torch
        self._device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self._dev_tri = {}       # id/bytes-keyed device copies of index arrays
        self._dev_by_tri = {}    # "v" / "c" / "n" -> CUDA tensor [T,3,3]
        self._ws = None

        array_vertices = np.array(vertices, dtype=np.float32)
        array_triangles_vertices = np.array(triangles_vertices, dtype=np.int32)
        if normals is not None and triangles_normals is not None:
            array_normals = np.array(normals, dtype=np.float32)
            array_triangles_normals = np.array(triangles_normals, dtype=np.int32)
        else:
            array_normals = None
            array_triangles_normals = None

        self._update_vertices_and_normals(array_vertices, array_triangles_vertices,
                                          array_normals, array_triangles_normals, recalculate_normals,
                                          invert_calculated_normals)

        if texture_coords is None or triangles_texture_coords is None or texture is None:
            self._texture_coords = None
            self._triangles_texture_coords = None
            self._texture = None
            self._colors = None
            self._colors_by_triangles = None
            self._dev_by_tri["c"] = None
        else:
            self._texture_coords = np.array(texture_coords, dtype=np.float32)
            self._triangles_texture_coords = np.array(triangles_texture_coords, dtype=np.int32)
            self._texture = np.array(texture)
            self._colors, self._colors_by_triangles = self._device_colors()

    def _stream(self):
        return self._torch.cuda.current_stream(self._device).cuda_stream

    def _to_device(self, a):
        t = self._torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self._device, non_blocking=False)

    def _gather(self, dev_attr, host_tri, n_rows, key):
        """attr[tri] on the device -> (host array [T,3,3], device tensor kept under `key`)."""
        torch = self._torch
        tri = _wrap_indices(host_tri, n_rows)
        if tri.size == 0:      # a mesh without faces: upstream's np.array([], int32) indexes to an empty (0,3) gather
            tri = tri.reshape(0, 3)
        if tri.ndim != 2 or tri.shape[1] != 3:
            raise ValueError(f"triangle index array must be [T,3], got {tuple(tri.shape)}")
        d_tri = self._to_device(tri)
        out = torch.empty((tri.shape[0], 3, 3), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            check(self._L.crb_model_gather(dev_attr.data_ptr(), d_tri.data_ptr(), tri.shape[0], out.data_ptr(),
                                           self._stream()))
        self._dev_by_tri[key] = out
        return out.cpu().numpy()

    def _device_colors(self):
        """model.py:146-151."""
        torch = self._torch
        tex = self._texture
        if tex.ndim != 3 or tex.shape[2] != 3 or tex.dtype != np.uint8:
            raise ValueError("texture must be a uint8 [h,w,3] image (what cv2.imread returns)")
        vt = self._texture_coords
        if vt.ndim != 2 or vt.shape[1] < 2:
            raise IndexError("index 1 is out of bounds for axis 1 with size %d" % (vt.shape[1] if vt.ndim == 2 else 0))
        h, w, _ = tex.shape
        d_vt, d_tex = self._to_device(vt), self._to_device(tex)
        d_col = torch.empty((vt.shape[0], 3), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            check(self._L.crb_model_vertex_colors(d_vt.data_ptr(), vt.shape[0], vt.shape[1], d_tex.data_ptr(), h, w,
                                                  d_col.data_ptr(), self._stream()))
        by_tri = self._gather(d_col, self._triangles_texture_coords, vt.shape[0], "c")
        return d_col.cpu().numpy(), by_tri

    def _device_normals(self, d_vertices, tri_wrapped, invert):
        """model.py:174-188 on the device -> CUDA tensor [V,3]."""
        torch = self._torch
        V, T = d_vertices.shape[0], tri_wrapped.shape[0]
        need = int(self._L.crb_model_normals_workspace_bytes(V, T))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self._device)
        d_tri = self._to_device(tri_wrapped)
        d_n = torch.empty((V, 3), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            check(self._L.crb_model_vertex_normals(d_vertices.data_ptr(), V, d_tri.data_ptr(), T, int(bool(invert)),
                                                   d_n.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                                   self._stream()))
        return d_n

    def _update_vertices_and_normals(self, array_vertices, array_triangles_vertices,
                                     array_normals, array_triangles_normals, recalculate_normals=True,
                                     invert_calculated_normals=False):
        """model.py:153-172."""
        self._vertices = array_vertices.astype('float32')
        self._triangles_vertices = array_triangles_vertices
        if self._vertices.ndim != 2 or self._vertices.shape[1] != 3:
            raise ValueError(f"vertices must be [V,3], got {tuple(self._vertices.shape)}")
        d_vertices = self._to_device(self._vertices)
        self._vertices_by_triangles = self._gather(d_vertices, self._triangles_vertices, len(self._vertices), "v")

        self._mean_vertex = self._vertices.mean(axis=0)
        self._max_span = np.max(np.linalg.norm(self._vertices - self._mean_vertex, axis=-1))

        if array_normals is not None and array_triangles_normals is not None and not recalculate_normals:
            self._normals = array_normals.astype('float32')
            self._triangles_normals = array_triangles_normals
            d_normals = self._to_device(self._normals)
        else:
            tri = _wrap_indices(self._triangles_vertices, len(self._vertices))
            d_normals = self._device_normals(d_vertices, tri, invert_calculated_normals)
            self._normals = d_normals.cpu().numpy()
            self._triangles_normals = self._triangles_vertices
        self._normals_by_triangles = self._gather(d_normals, self._triangles_normals, len(self._normals), "n")

    # ------------------------------------------------------------------------------------------------ access
    def device_triangles(self):
        """(vertices, colours, normals) by triangle as CUDA float32 tensors [T,3,3] -- the device-resident twins of
        `_vertices_by_triangles / _colors_by_triangles / _normals_by_triangles` (for `filler.render_arrays`)."""
        return self._dev_by_tri["v"], self._dev_by_tri.get("c"), self._dev_by_tri["n"]

    def get_vertex(self, index: int):
        return (self._vertices[index], (self._colors[index] if self._colors is not None else None),
                self._normals[index])

    def get_triangle(self, index: int):
        return (self._vertices_by_triangles[index],
                (self._colors_by_triangles[index] if self._colors_by_triangles is not None else None),
                self._normals_by_triangles[index])

    def n_triangles(self) -> int:
        return len(self._triangles_vertices)

    def n_vertices(self) -> int:
        return len(self._vertices)

    def get_mean_vertex(self):
        return self._mean_vertex

    def get_max_span(self):
        return self._max_span

    # ------------------------------------------------------------------------------------------------ transforms
    def shift(self, shift):
        """model.py:213-216: normals are kept."""
        moved = self._vertices + shift
        self._update_vertices_and_normals(moved, self._triangles_vertices, self._normals, self._triangles_normals,
                                          recalculate_normals=False)

    def scale(self, scale_coef, keep_position=True):
        """model.py:218-227: in place on the float32 vertex array, about the mean vertex unless told otherwise."""
        scaled = self._vertices
        if keep_position:
            scaled -= self._mean_vertex
            scaled *= scale_coef
            scaled += self._mean_vertex
        else:
            scaled *= scale_coef
        self._update_vertices_and_normals(scaled, self._triangles_vertices, self._normals, self._triangles_normals,
                                          recalculate_normals=False)

    @staticmethod
    def _rot_matrix(angle, degrees=True):
        """model.py:229-236: [[c, s], [-s, c]] in float64."""
        if degrees:
            angle *= np.pi / 180
        c, s = np.cos(angle), np.sin(angle)
        return np.array([[c, s], [-s, c]])

    def rotate(self, angles):
        """model.py:238-256: v @ (Rx Ry Rz)^T in float64, narrowed by the update; normals recomputed (on the device)."""
        assert len(angles) == 3
        rx, ry, rz = np.eye(3), np.eye(3), np.eye(3)
        rx[1:, 1:] = Model._rot_matrix(angles[0])
        ry[::2, ::2] = Model._rot_matrix(angles[1])
        rz[:2, :2] = Model._rot_matrix(angles[2])
        rot = np.matmul(np.matmul(rx, ry), rz)
        turned = np.matmul(self._vertices, np.transpose(rot))
        self._update_vertices_and_normals(turned, self._triangles_vertices, None, None, recalculate_normals=True)
