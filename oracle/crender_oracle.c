/*
 * crender_oracle.c -- CPU restatement of the reference's Version C hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product package (cython3dmodelrenderer_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_pinning.py checks this file bit-for-bit against the
 * reference's own Cython build (oracle/_ref, built by oracle/build_ref.py from /root/reference) wherever
 * that build is present, and against tests/golden/ checksums that were produced by that build
 * (tests/golden/make_golden.py) everywhere else.
 *
 * Every function cites the reference lines it restates.  `pyx` below means
 *   crender/cy/pixel_buffer_filler/advanced_pixel_buffer_filler.pyx
 * and `mu` means crender/cy/pixel_buffer_filler/math_utils.pyx (both relative to the reference root).
 *
 * All arithmetic is IEEE binary32, one rounding per operation, no fused multiply-add: build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math   (x86-64 SSE2: FLT_EVAL_METHOD == 0).
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- a1: constructor scalars and projection matrix --------------------------------------------------
 * pyx:54-60  fov/z_near/z_far are stored as C floats; f = 1/tan(fov/2/180*pi) is evaluated in double
 *            on the float fov and then narrowed; a = h/w is a Python true division narrowed to float.
 * pyx:83-90  q = z_far/(z_far-z_near) in float; P[0][0] = f/a and P[3][2] = -z_near*q in float.
 * Returns 0, or -1 where the reference raises ZeroDivisionError (w == 0, h == 0, z_far == z_near).
 * proj is row-major 4x4, the layout np.array(..., dtype='float32') gives it.                            */
int oracle_projection(int h, int w, float fov, float z_near, float z_far, float proj[16])
{
    if (w == 0) return -1;
    float a = (float)((double)h / (double)w);
    double ang = (((double)fov / 2.0) / 180.0) * M_PI;
    float f = (float)(1.0 / tan(ang));
    float d = z_far - z_near;
    if (d == 0.0f || a == 0.0f) return -1;
    float q = z_far / d;
    memset(proj, 0, 16 * sizeof(float));
    proj[0] = f / a;
    proj[5] = f;
    proj[10] = q;
    proj[11] = 1.0f;
    proj[14] = (-z_near) * q;
    return 0;
}

/* ---- a3: in-place projection of one vertex ------------------------------------------------------------
 * pyx:116-130.  The reference calls project_on_screen_multithread(triangles, triangles): source and
 * destination alias, so column j reads the components columns <j already overwrote.  With the zero
 * pattern of P those stale reads are multiplied by 0, but they are still evaluated, which matters for
 * inf/NaN inputs -- hence the literal restatement.  x_scale = (float)(w/2.0), y_scale = (float)(h/2.0).  */
void oracle_project_vertex(const float P[16], int h, int w, float vert[3])
{
    float x_scale = (float)((double)w / 2.0), y_scale = (float)((double)h / 2.0);
    float z = vert[2];
    for (int j = 0; j < 3; ++j)
        vert[j] = vert[0] * P[0 * 4 + j] + vert[1] * P[1 * 4 + j] + vert[2] * P[2 * 4 + j] + P[3 * 4 + j];
    vert[0] = vert[0] / z;
    vert[1] = vert[1] / z;
    vert[2] = vert[2] / z;
    vert[0] = vert[0] + 1.0f;
    vert[1] = vert[1] + 1.0f;
    vert[0] = vert[0] * x_scale;
    vert[1] = vert[1] * y_scale;
}

/* (int)ceil(x) as the reference's compiled code performs it (pyx:165-166).  The C cast is undefined for
 * values outside int; the reference binary (gcc, x86-64) uses cvttsd2si, which returns INT_MIN for
 * every out-of-range or non-finite input.  Stated here explicitly so the GPU path can match it.         */
static int ceil_to_int(float x)
{
    double c = ceil((double)x);
    if (!(c > -2147483649.0 && c < 2147483648.0)) return INT_MIN;
    return (int)c;
}

static int clip_int(int a, int lo, int hi) /* math_utils.pxd:8-13 */
{
    if (a < lo) return lo;
    if (a > hi) return hi;
    return a;
}

/* ---- a5: half-open pixel rectangle of a projected triangle -- pyx:132-175 ------------------------------
 * out = {x_left, x_right, y_top, y_bot}; running minima start at (w, h), maxima at 0; NaNs never win a
 * comparison.  Pixel centres sit on integer coordinates.                                                 */
void oracle_pixel_rect(const float tri[9], int h, int w, int out[4])
{
    float xl = (float)w, xr = 0.0f, yt = (float)h, yb = 0.0f;
    for (int i = 0; i < 3; ++i) {
        float x = tri[i * 3], y = tri[i * 3 + 1];
        if (x < xl) xl = x;
        if (x > xr) xr = x;
        if (y < yt) yt = y;
        if (y > yb) yb = y;
    }
    out[0] = clip_int(ceil_to_int(xl), 0, w);
    out[1] = clip_int(ceil_to_int(xr), 0, w);
    out[2] = clip_int(ceil_to_int(yt), 0, h);
    out[3] = clip_int(ceil_to_int(yb), 0, h);
}

/* ---- a6: barycentric coordinates of pixel (x, y) -- mu:5-34 ---------------------------------------------
 * Each coordinate is (l1*(py-a) - l2*(px-b))/l3 with a true float division (cdivision: x/0 -> inf/NaN). */
void oracle_barycentric(const float tri[9], int x, int y, float bar[3])
{
    float x0 = tri[0], y0 = tri[1], x1 = tri[3], y1 = tri[4], x2 = tri[6], y2 = tri[7];
    float px = (float)x, py = (float)y;
    float l01 = x1 - x2, l02 = y1 - y2;
    float l03 = l01 * (y0 - y2) - l02 * (x0 - x2);
    float l11 = x2 - x0, l12 = y2 - y0;
    float l13 = l11 * (y1 - y0) - l12 * (x1 - x0);
    float l21 = x0 - x1, l22 = y0 - y1;
    float l23 = l21 * (y2 - y1) - l22 * (x2 - x1);
    bar[0] = (l01 * (py - y2) - l02 * (px - x2)) / l03;
    bar[1] = (l11 * (py - y0) - l12 * (px - x0)) / l13;
    bar[2] = (l21 * (py - y1) - l22 * (px - x1)) / l23;
}

/* ---- a4 + a7: draw one projected triangle into rows [row0, row1) -- pyx:202-242 --------------------------
 * tri = projected vertices, nrm/col = the triangle's 3x3 normals / colours.  Sequential semantics of the
 * reference at n_threads=1: a fragment is dropped when any barycentric < 0, when its depth is NaN
 * (pyx:220 is a tautology for every other value) or when depth > z_buffer; an equal depth overwrites.   */
static void draw_triangle(const float tri[9], const float col[9], const float nrm[9], int h, int w,
                          int row0, int row1, float *zbuf, float *cbuf, float *nbuf)
{
    if ((double)(nrm[2] + nrm[5] + nrm[8]) / 3.0 >= 0.0) return; /* pyx:202-204 */
    int r[4];
    oracle_pixel_rect(tri, h, w, r);
    if (r[0] - r[1] == 0 || r[2] - r[3] == 0) return; /* pyx:209-211 */
    int y_lo = r[2] > row0 ? r[2] : row0, y_hi = r[3] < row1 ? r[3] : row1;
    for (int x = r[0]; x < r[1]; ++x) {
        for (int y = y_lo; y < y_hi; ++y) {
            float b[3];
            oracle_barycentric(tri, x, y, b);
            if (b[0] < 0.0f || b[1] < 0.0f || b[2] < 0.0f) continue;
            float new_z = tri[2] * b[0] + tri[5] * b[1] + tri[8] * b[2];
            if (!(-1.0 <= new_z || new_z <= 1.0)) continue;
            size_t p = (size_t)y * (size_t)w + (size_t)x;
            if (new_z > zbuf[p]) continue;
            float n0 = nrm[0] * b[0] + nrm[3] * b[1] + nrm[6] * b[2];
            float n1 = nrm[1] * b[0] + nrm[4] * b[1] + nrm[7] * b[2];
            float n2 = nrm[2] * b[0] + nrm[5] * b[1] + nrm[8] * b[2];
            float c0 = col[0] * b[0] + col[3] * b[1] + col[6] * b[2];
            float c1 = col[1] * b[0] + col[4] * b[1] + col[7] * b[2];
            float c2 = col[2] * b[0] + col[5] * b[1] + col[8] * b[2];
            zbuf[p] = new_z;
            cbuf[p * 3 + 0] = c0; cbuf[p * 3 + 1] = c1; cbuf[p * 3 + 2] = c2;
            nbuf[p * 3 + 0] = n0; nbuf[p * 3 + 1] = n1; nbuf[p * 3 + 2] = n2;
        }
    }
}

/* ---- a2: render_model -- pyx:92-104 ----------------------------------------------------------------------
 * v/c/n: [T,3,3] float32, C-contiguous, never modified (the reference works on .copy()s).  zbuf [h,w],
 * cbuf/nbuf [h,w,3] are the filler's persistent buffers and are composited into, not cleared.
 * n_threads <= 1: the reference's deterministic single-thread order (triangle index ascending).
 * n_threads  > 1: rows are split into n_threads bands, each band walks all triangles in index order and
 *                 touches only its own rows -- bit-identical to the single-thread result by construction
 *                 (used only to time a many-core CPU baseline when oracle/_ref is unavailable).
 * Returns 0, or -2 if scratch memory cannot be allocated.                                               */
int oracle_render(int h, int w, const float P[16], const float *v, const float *c, const float *n,
                  int64_t T, float *zbuf, float *cbuf, float *nbuf, int n_threads)
{
    float *scr = (float *)malloc((size_t)(T > 0 ? T : 1) * 9 * sizeof(float));
    if (!scr) return -2;
    memcpy(scr, v, (size_t)T * 9 * sizeof(float));
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 1 ? n_threads : 1)
#endif
    for (int64_t i = 0; i < T; ++i)
        for (int k = 0; k < 3; ++k) oracle_project_vertex(P, h, w, scr + i * 9 + k * 3);

    if (n_threads <= 1) {
        for (int64_t i = 0; i < T; ++i)
            draw_triangle(scr + i * 9, c + i * 9, n + i * 9, h, w, 0, h, zbuf, cbuf, nbuf);
    } else {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
        for (int b = 0; b < n_threads * 8; ++b) {
            int row0 = (int)((int64_t)h * b / (n_threads * 8)), row1 = (int)((int64_t)h * (b + 1) / (n_threads * 8));
            for (int64_t i = 0; i < T; ++i)
                draw_triangle(scr + i * 9, c + i * 9, n + i * 9, h, w, row0, row1, zbuf, cbuf, nbuf);
        }
    }
    free(scr);
    return 0;
}

/* Project only: writes the [T,3,3] screen-space triangles the raster stage consumes (test helper). */
void oracle_project(int h, int w, const float P[16], const float *v, int64_t T, float *out)
{
    memcpy(out, v, (size_t)T * 9 * sizeof(float));
    for (int64_t i = 0; i < T * 3; ++i) oracle_project_vertex(P, h, w, out + i * 3);
}

/* Fresh-filler buffers -- pyx:65-67: normals 0, colour 0, z = ones*1e6 (float32 1e6). */
void oracle_init_buffers(int h, int w, float *zbuf, float *cbuf, float *nbuf)
{
    size_t px = (size_t)h * (size_t)w;
    for (size_t i = 0; i < px; ++i) zbuf[i] = 1e6f;
    memset(cbuf, 0, px * 3 * sizeof(float));
    memset(nbuf, 0, px * 3 * sizeof(float));
}

/* ---- N1: GuroIllumination.draw_illumination -- crender/cy/illumination/guro_illumination.py:20-27 --------
 * NumPy float32 semantics restated per pixel: dot = n0*l0 + n1*l1 + n2*l2 (np.sum over 3 terms, pairwise
 * == left-to-right for n < 8), norm = sqrt(n0^2+n1^2+n2^2) (np.linalg.norm: sqrt of sum of squares in f32),
 * shadow = clip(dot / (norm + 1e-6), 0, 1) (1e-6 is a Python float: NumPy keeps float32), color *= shadow. */
void oracle_guro(int h, int w, const float light[3], float *cbuf, const float *nbuf)
{
    size_t px = (size_t)h * (size_t)w;
    for (size_t i = 0; i < px; ++i) {
        const float *nn = nbuf + i * 3;
        /* np.sum(n_buffer * light, axis=-1) (guro_illumination.py:23) accumulates from the identity +0.0, so products
         * that are all -0.0 (zero normal components times a light of (-0,-0,-1), the default) sum to +0.0, not -0.0:
         * checked against NumPy 2.3 through the golden lit-colour checksums */
        float dot = 0.0f;
        dot += nn[0] * light[0];
        dot += nn[1] * light[1];
        dot += nn[2] * light[2];
        float nrm = sqrtf(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
        float s = dot / (nrm + 1e-6f);
        if (s < 0.0f) s = 0.0f;
        if (s > 1.0f) s = 1.0f;
        cbuf[i * 3 + 0] *= s; cbuf[i * 3 + 1] *= s; cbuf[i * 3 + 2] *= s;
    }
}
