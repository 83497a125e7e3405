/*
 * ingest_oracle.c -- CPU restatement of the reference's model ingest arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load
 * this.  The product package (cython3dmodelrenderer_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_ingest_oracle.py checks every function bit-for-bit against the
 * reference's own `Model` (crender/cy/data_structures/model.py, imported from oracle/_ref where present) and
 * against tests/golden/ingest_*.npz, which that class produced in the build container
 * (tests/golden/make_golden_ingest.py).
 *
 * `model.py` below is crender/cy/data_structures/model.py relative to the reference root.  The reference is
 * NumPy code, so its rounding is NumPy's (2.3.5 + OpenBLAS 0.3.30, the versions of this image):
 *   - elementwise float32 ufuncs round once per operation, no FMA;
 *   - a float32 `dot` of two 3-vectors (np.dot, and np.linalg.norm of a 1-D vector = sqrt(x.dot(x))) goes to
 *     cblas_sdot, whose x86-64 kernel sums the tail elements (all of them when n < 32) as float32 products in a
 *     DOUBLE accumulator and narrows once at the end  [probed: 100 % of 20 000 random pairs; a float accumulator
 *     in any order matches 78 %];
 *   - np.mean(stack, axis=0) adds the rows in order in float32, starting from +0.0, and divides by the count (the division is done
 *     in double on the float sum and narrowed; with 53 >= 2*24+2 bits that equals the float32 quotient).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* np.dot of float32 3-vectors (see header). */
static float dot3(const float *a, const float *b)
{
    float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2];
    double acc = 0.0;
    acc += (double)p0;
    acc += (double)p1;
    acc += (double)p2;
    return (float)acc;
}

/* model.py:190-194 _normalize: n / norm(n) unless norm(n) == 0 (NaN norms divide). */
static void normalize3(float *n)
{
    float nrm = sqrtf(dot3(n, n));
    if (nrm == 0.0f) return;
    n[0] = n[0] / nrm;
    n[1] = n[1] / nrm;
    n[2] = n[2] / nrm;
}

/* model.py:196-201 _compute_triangle_normal: -cross(t1 - t0, t1 - t2), normalised.  np.cross forms each
 * component as (a_i * b_j) - (a_j * b_i) with both products rounded to float32. */
void ingest_face_normal(const float *t0, const float *t1, const float *t2, float n[3])
{
    float a[3], b[3];
    for (int k = 0; k < 3; ++k) {
        a[k] = t1[k] - t0[k];
        b[k] = t1[k] - t2[k];
    }
    float c0 = a[1] * b[2];
    c0 -= a[2] * b[1];
    float c1 = a[2] * b[0];
    c1 -= a[0] * b[2];
    float c2 = a[0] * b[1];
    c2 -= a[1] * b[0];
    n[0] = -c0;
    n[1] = -c1;
    n[2] = -c2;
    normalize3(n);
}

typedef struct {
    float *n;  /* kept face normals of one vertex, in the order they were met */
    int count, cap;
} nlist;

/* model.py:174-188 _compute_normals_by_vertex(vertices [V,3], triangles [T,3]) with
 * duplicate_normal_dot_tolerance = 0; `invert` = model.py:168-169.  Triangle indices must already be in
 * [0,V) (NumPy and the Python list both wrap negative ones; the caller does that).
 * For every (triangle, corner) in order: the face normal joins the corner vertex's list unless some normal
 * already in the list has dot >= 1 with it.  Result per vertex: normalise(mean of its list), or zeros. */
int ingest_vertex_normals(const float *vertices, int64_t V, const int32_t *tri, int64_t T, int invert, float *out)
{
    nlist *lists = (nlist *)calloc((size_t)(V > 0 ? V : 1), sizeof(nlist));
    if (!lists) return -1;
    for (int64_t t = 0; t < T; ++t) {
        float n[3];
        ingest_face_normal(vertices + 3 * (int64_t)tri[3 * t], vertices + 3 * (int64_t)tri[3 * t + 1],
                           vertices + 3 * (int64_t)tri[3 * t + 2], n);
        for (int c = 0; c < 3; ++c) {
            nlist *L = &lists[tri[3 * t + c]];
            int is_new = 1;
            for (int j = 0; j < L->count; ++j)
                if (dot3(L->n + 3 * j, n) >= 1.0f) is_new = 0;
            if (!is_new) continue;
            if (L->count == L->cap) {
                L->cap = L->cap ? 2 * L->cap : 8;
                L->n = (float *)realloc(L->n, sizeof(float) * 3 * (size_t)L->cap);
                if (!L->n) return -1;
            }
            memcpy(L->n + 3 * L->count, n, sizeof n);
            L->count++;
        }
    }
    for (int64_t v = 0; v < V; ++v) {
        nlist *L = &lists[v];
        float m[3] = {0.0f, 0.0f, 0.0f};
        if (L->count > 0) {
            /* np.mean(axis=0): add.reduce starts from the identity +0.0 (so a lone -0.0 becomes +0.0  [probed])
             * and adds the rows in order */
            for (int j = 0; j < L->count; ++j)
                for (int k = 0; k < 3; ++k) m[k] = m[k] + L->n[3 * j + k];
            for (int k = 0; k < 3; ++k) m[k] = (float)((double)m[k] / (double)L->count);
            normalize3(m);
        }
        if (invert)
            for (int k = 0; k < 3; ++k) m[k] = m[k] * -1.0f;
        memcpy(out + 3 * v, m, sizeof m);
        free(L->n);
    }
    free(lists);
    return 0;
}

/* (int32) cast of a float the way NumPy's astype does it on x86-64: cvttss2si, which yields INT32_MIN for
 * NaN and for anything outside the int32 range. */
static int32_t f32_to_i32(float x)
{
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)x;
}

static int clipi(int v, int lo, int hi)
{
    /* np.clip(v, lo, hi) = minimum(maximum(v, lo), hi) */
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

/* model.py:147-150: colours[i] = texture[clip(int32((1 - vt[i,1]) * h), 0, h-1), clip(int32(vt[i,0] * w), 0, w-1)]
 * as float32; texture is uint8 [h,w,3] (cv2.imread, BGR); vt rows have `width` (2 or 3) floats. */
void ingest_vertex_colors(const float *vt, int64_t n, int width, const uint8_t *texture, int h, int w, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        float fy = (1.0f - vt[width * i + 1]) * (float)h;
        float fx = vt[width * i] * (float)w;
        int y = clipi(f32_to_i32(fy), 0, h - 1), x = clipi(f32_to_i32(fx), 0, w - 1);
        const uint8_t *px = texture + 3 * ((int64_t)y * w + x);
        for (int k = 0; k < 3; ++k) out[3 * i + k] = (float)px[k];
    }
}

/* model.py:151,158,172: attr[tri] fancy-index gather, [n,3] x [T,3] -> [T,3,3]; indices already in [0,n). */
void ingest_gather(const float *attr, const int32_t *tri, int64_t T, float *out)
{
    for (int64_t e = 0; e < 3 * T; ++e) memcpy(out + 3 * e, attr + 3 * (int64_t)tri[e], 3 * sizeof(float));
}
