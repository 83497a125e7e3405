/*
 * crender_ingest_b200.h -- C ABI of the model-ingest row of libcrender_b200.so: what stands
 * between an .obj file and the three [T,3,3] arrays the rendering hot path (crender_b200.h) reads.
 *
 * The reference does this in Python: `Model.read_model` / `Model.__init__` / `Model.rotate`
 * (crender/cy/data_structures/model.py, "model.py" below).  There is no FFI upstream; a maintainer would bind these
 * entry points from model.py with ctypes (INTEGRATION.md shows the stub).  Its cost today: 0.7 s (T-Rex) to 2.7 s
 * (bunny) per load and ~0.7 s per `rotate`, nearly all of it the per-triangle Python loop of
 * `_compute_normals_by_vertex` (model.py:174-188).
 *
 * Conventions are those of crender_b200.h: plain C types, CRB_OK or a negative CRB_ERR_* code, crb_last_error() for
 * the text.  "DEVICE" pointers live on the calling thread's current CUDA device; kernels are queued on `stream`
 * (cudaStream_t as void*) and the calls return without waiting for them.
 *
 * Arithmetic is NumPy's, restated: float32, one rounding per operation, no FMA; a 3-element float32 dot product
 * (np.dot, np.linalg.norm) = the three float32 products added in a double accumulator and narrowed once, which is
 * what cblas_sdot does for short vectors; np.mean = ordered float32 sum from +0.0, divided by the count.
 */
#ifndef CRENDER_INGEST_B200_H
#define CRENDER_INGEST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define CRB_ERR_SYNTAX (-6)   /* an .obj token Python would accept but this reader does not (digit-group underscores,
                                 non-ASCII digits): refused loudly rather than read differently */
#define CRB_ERR_RANGE (-7)    /* a face index outside int32 (np.array(..., dtype=int32) raises OverflowError) */

/* ---- .obj reader (HOST): model.py:7-77 read_model, 258-312 _read_vertex/_read_texture_coord/_read_normal/
 * _fix_index/_read_face ----------------------------------------------------------------------------------------
 * Same line discipline as the reference: universal newlines; a line whose first character is '#' is a comment;
 * the command is the text before the FIRST space (so "v\t1 2 3" is not a vertex); `v` needs >= 3 floats (first 3
 * kept), `vn` exactly 3, `vt` keeps all of its floats; `f` is fan-triangulated (c0, c[i+1], c[i+2]) with corners
 * "v", "v/vt", "v//vn", "v/vt/vn", indices made 0-based when positive and kept as they are when 0 or negative; the
 * per-triangle vt / vn lists stop for good at the first face lacking them; a line on which Python would raise is
 * skipped whole (silent=True).  `mtllib` payloads are kept in order for the caller (texture lookup stays in Python:
 * cv2.imread, as upstream). */
typedef struct crb_obj crb_obj;

int crb_obj_parse(const char *text, size_t bytes, crb_obj **out);
void crb_obj_free(crb_obj *o);

/* counts[10] = { vertices, texture coords, floats per texture coord (0 if none; -1 if rows differ in length -- the
 * reference's np.array then raises ValueError), normals, triangles, 1 if every face carried vt, 1 if every face
 * carried vn, mtllib lines, lines Python would have raised on (skipped), upstream's `line_index + 1` of the first
 * such line (what silent=False reports) or 0 } */
int crb_obj_counts(const crb_obj *o, int64_t counts[10]);

/* Copies into caller arrays (any may be NULL): v [n_v,3], vt [n_vt,width], vn [n_vn,3] float32 (the float32 narrowing of
 * Python's float(), as np.array(list, dtype=float32) does), tri_v / tri_vt / tri_vn [n_tri,3] int32. */
int crb_obj_copy(const crb_obj *o, float *v, float *vt, float *vn, int32_t *tri_v, int32_t *tri_vt, int32_t *tri_vn);

/* k-th `mtllib` payload (not NUL-terminated; valid until crb_obj_free). */
int crb_obj_mtllib(const crb_obj *o, int k, const char **data, size_t *len);

/* ---- smooth vertex normals (DEVICE): model.py:174-188 _compute_normals_by_vertex, 190-201, 168-169 ------------
 * vertices [V,3] f32, tri [T,3] int32 with every index already in [0,V) (NumPy wraps negative ones; the host does
 * that before upload).  For every (triangle, corner) in file order the triangle's unit normal
 * -cross(t1-t0, t1-t2)/|.| joins the corner vertex's list unless a normal already in the list has dot >= 1 with it;
 * normals_out[v] = normalise(mean(list)), zeros for an unreferenced vertex, times -1 if `invert`.
 * Workspace: crb_model_normals_workspace_bytes(V,T) bytes of DEVICE scratch (about 76 B/triangle + 12 B/vertex). */
size_t crb_model_normals_workspace_bytes(int64_t V, int64_t T);
int crb_model_vertex_normals(const float *vertices, int64_t V, const int32_t *tri, int64_t T, int invert,
                             float *normals_out, void *workspace, size_t workspace_bytes, void *stream);

/* ---- per-vertex texture colours (DEVICE): model.py:147-150 ------------------------------------------------------
 * colors_out[i] = (float) texture[clip((int32)((1 - vt[i,1]) * h), 0, h-1)][clip((int32)(vt[i,0] * w), 0, w-1)],
 * texture uint8 [h,w,3] (BGR as cv2.imread returns it), vt [n,width] f32, width >= 2; out-of-range casts give
 * INT32_MIN like x86's cvttss2si. */
int crb_model_vertex_colors(const float *vt, int64_t n, int width, const uint8_t *texture, int tex_h, int tex_w,
                            float *colors_out, void *stream);

/* ---- by-triangle gather (DEVICE): model.py:151,158,172  attr[tri] ----------------------------------------------
 * attr [n,3] f32, tri [T,3] int32 in [0,n) -> out [T,3,3] f32: one of the three arrays crb_render reads. */
int crb_model_gather(const float *attr, const int32_t *tri, int64_t T, float *out, void *stream);

/* Kernel launches issued by the crb_model_* entry points of this process (bench / tests: proof the device path ran). */
int64_t crb_model_launch_count(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* CRENDER_INGEST_B200_H */
