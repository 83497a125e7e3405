// property test of the closed-form conservative span (no walk) against the per-pixel threshold test
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
static uint64_t rs = 88172645463325252ull;
static uint64_t rnd(void){ rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }
static double urand(void){ return (rnd() >> 11) * (1.0 / 9007199254740992.0); }
static float rcp_approx(float x){ // <= 1 ulp error model, ftz
    if (fabsf(x) < 1.17549435e-38f) return copysignf(INFINITY, x);
    float r = (float)(1.0 / (double)x);
    int k = (int)(rnd() % 3) - 1;  // -1,0,+1 ulp
    uint32_t b; __builtin_memcpy(&b, &r, 4); b += k; __builtin_memcpy(&r, &b, 4);
    return r;
}
static inline float fminf_nan(float a, float b){ return (b != b) ? a : (a != a ? b : (a < b ? a : b)); }
static inline float fmaxf_nan(float a, float b){ return (b != b) ? a : (a != a ? b : (a > b ? a : b)); }
#define REJ_EPS 1e-6f
static void edge_bound(float A, float l2, float xb, float xc, float *lo, float *hi)
{
    if (!(fabsf(l2) >= 1e-20f)) return;
    const float w = fabsf(xc - xb) + 16.0f;
    const float M = fmaf(fmaf(fabsf(l2), w, fabsf(A)), 3.814697e-6f, 1e-5f);
    const float s = A + M;
    const float d = s * rcp_approx(l2);
    const float e = d + xb;
    const float wd = fmaf(fabsf(d), 9.5367431640625e-7f, 0.0078125f);
    if (l2 > 0.0f) *hi = fminf_nan(*hi, e + wd); else *lo = fmaxf_nan(*lo, e - wd);
}
int main(int argc, char **argv)
{
    long N = argc > 1 ? atol(argv[1]) : 2000000;
    long bad = 0, rows = 0, frags = 0, cand = 0, extra = 0;
    for (long it = 0; it < N; ++it) {
        // random triangle in screen space
        int mode = it % 6;
        float scale = mode == 0 ? 5.f : mode == 1 ? 40.f : mode == 2 ? 400.f : mode == 3 ? 3000.f : mode == 4 ? 1.5f : 200000.f;
        float cx = (float)(urand() * 4096), cy = (float)(urand() * 4096);
        float x[3], y[3];
        for (int k = 0; k < 3; ++k) { x[k] = cx + (float)((urand() * 2 - 1) * scale); y[k] = cy + (float)((urand() * 2 - 1) * scale); }
        if (it % 11 == 0) { y[1] = y[0]; }              // horizontal edge
        if (it % 13 == 0) { x[2] = x[1]; }              // vertical edge
        if (it % 17 == 0) { x[0] = roundf(x[0]); y[0] = roundf(y[0]); x[1] = roundf(x[1]); y[1] = roundf(y[1]); }
        if (it % 19 == 0) { x[2] = x[0] + (x[1]-x[0])*0.5f; y[2] = y[0] + (y[1]-y[0])*0.5f + 1e-3f; } // sliver
        const float l03 = (x[1] - x[2]) * (y[0] - y[2]) - (y[1] - y[2]) * (x[0] - x[2]);
        const float l13 = (x[2] - x[0]) * (y[1] - y[0]) - (y[2] - y[0]) * (x[1] - x[0]);
        const float l23 = (x[0] - x[1]) * (y[2] - y[1]) - (y[0] - y[1]) * (x[2] - x[1]);
        if (!(fabsf(l03) >= 1e-30f && fabsf(l13) >= 1e-30f && fabsf(l23) >= 1e-30f)) continue;
        float s1 = l03 < 0 ? -1.f : 1.f, s2 = l13 < 0 ? -1.f : 1.f, s3 = l23 < 0 ? -1.f : 1.f;
        float l01 = s1 * (x[1] - x[2]), l02 = s1 * (y[1] - y[2]);
        float l11 = s2 * (x[2] - x[0]), l12 = s2 * (y[2] - y[0]);
        float l21 = s3 * (x[0] - x[1]), l22 = s3 * (y[0] - y[1]);
        // a tile: choose around the triangle
        int tx0 = ((int)fmaxf(0.f, fminf(x[0], fminf(x[1], x[2])) + (float)(urand() * 40 - 10)) / 32) * 32;
        float ymin = fminf(y[0], fminf(y[1], y[2])), ymax = fmaxf(y[0], fmaxf(y[1], y[2]));
        float xmin = fminf(x[0], fminf(x[1], x[2])), xmax = fmaxf(x[0], fmaxf(x[1], x[2]));
        int bxl = (int)ceilf(fmaxf(xmin, 0)), bxr = (int)ceilf(fminf(xmax, 65535.f));
        int xa = bxl > tx0 ? bxl : tx0, xb = bxr < tx0 + 32 ? bxr : tx0 + 32;
        if (xa >= xb) continue;
        int yy = (int)ceilf(ymin) + (int)(urand() * (ymax - ymin + 1));
        if (yy < 0) yy = 0;
        const float py = (float)yy;
        const float A1 = l01 * (py - y[2]), A2 = l11 * (py - y[0]), A3 = l21 * (py - y[1]);
        float lo = (float)xa, hi = (float)(xb - 1);
        const float xc = (float)(tx0 + 16);
        edge_bound(A1, l02, x[2], xc, &lo, &hi);
        edge_bound(A2, l12, x[0], xc, &lo, &hi);
        edge_bound(A3, l22, x[1], xc, &lo, &hi);
        int sa = xa, sb = xb;
        if (lo <= hi) { int a = (int)ceilf(lo), b = (int)floorf(hi) + 1; if (a > sa) sa = a; if (b < sb) sb = b; } else sb = sa;
        ++rows;
        int first = -1, last = -1;
        for (int px = xa; px < xb; ++px) {
            const float fx = (float)px;
            const float n1 = A1 - l02 * (fx - x[2]), n2 = A2 - l12 * (fx - x[0]), n3 = A3 - l22 * (fx - x[1]);
            const int keep = !(n1 < -REJ_EPS || n2 < -REJ_EPS || n3 < -REJ_EPS);
            ++cand;
            if (keep) { ++frags; if (first < 0) first = px; last = px; if (px < sa || px >= sb) { if (bad < 10) printf("BAD it=%ld px=%d span=[%d,%d) rect=[%d,%d) lo=%g hi=%g\n", it, px, sa, sb, xa, xb, lo, hi); ++bad; } }
        }
        if (sb > sa) extra += (sb - sa) - (first < 0 ? 0 : last - first + 1);
    }
    printf("rows %ld candidates %ld survivors %ld extra-in-span %ld (%.3f%% of survivors) BAD %ld\n", rows, cand, frags, extra, 100.0 * extra / (frags ? frags : 1), bad);
    return bad != 0;
}
