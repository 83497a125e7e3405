// ingest_b200.cu -- model ingest of libcrender_b200.so: the .obj reader (host) and the
// kernels that turn indexed mesh data into the [T,3,3] arrays the rendering path reads (sm_100a).
// C ABI: include/crender_ingest_b200.h.  "model.py" = crender/cy/data_structures/model.py of the reference.
//
// Built with -fmad=false like the rest of the library: NumPy rounds after every multiply and add.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "crender_b200.h"
#include "crender_ingest_b200.h"

int crb_internal_fail(int code, const char *msg);   // crender_b200.cu: sets the calling thread's crb_last_error text

namespace {

int failf(int code, const char *fmt, ...)
{
    char buf[400];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return crb_internal_fail(code, buf);
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return failf(CRB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

std::atomic<int64_t> g_launches{0};

// =============================================================================================================
// .obj reader (host)
// =============================================================================================================

// str.split() separators that can occur in a line of an ASCII file (the newline characters are gone by then)
inline bool is_space(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f); }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

enum Tok { TOK_OK, TOK_BAD /* Python raises ValueError: the line is skipped */, TOK_REFUSE /* see CRB_ERR_SYNTAX */ };

// A token Python might read differently from the plain grammar below: digit-group underscores ("1_000") and
// non-ASCII characters (full-width digits are digits to float()/int()).
inline bool exotic(const char *p, const char *e)
{
    for (; p < e; ++p)
        if (*p == '_' || (unsigned char)*p >= 0x80) return true;
    return false;
}

bool word_is(const char *p, const char *e, const char *w)
{
    size_t n = strlen(w);
    if ((size_t)(e - p) != n) return false;
    for (size_t i = 0; i < n; ++i)
        if ((p[i] | 0x20) != w[i]) return false;
    return true;
}

// float(token): [+-]? ( digits [. digits*] | . digits ) ( [eE] [+-]? digits )?  |  [+-]? inf | infinity | nan
Tok read_float(const char *p, const char *e, double *out)
{
    const char *q = p;
    bool neg = false;
    if (q < e && (*q == '+' || *q == '-')) neg = (*q++ == '-');
    if (word_is(q, e, "inf") || word_is(q, e, "infinity")) {
        *out = neg ? -INFINITY : INFINITY;
        return TOK_OK;
    }
    if (word_is(q, e, "nan")) {
        *out = neg ? -NAN : NAN;
        return TOK_OK;
    }
    const char *s = q;
    int digits = 0;
    while (s < e && is_digit(*s)) ++s, ++digits;
    if (s < e && *s == '.') {
        ++s;
        while (s < e && is_digit(*s)) ++s, ++digits;
    }
    bool ok = digits > 0;
    if (ok && s < e && (*s == 'e' || *s == 'E')) {
        ++s;
        if (s < e && (*s == '+' || *s == '-')) ++s;
        int ed = 0;
        while (s < e && is_digit(*s)) ++s, ++ed;
        ok = ed > 0;
    }
    if (!ok || s != e) return exotic(p, e) ? TOK_REFUSE : TOK_BAD;
    double v = 0.0;
    auto r = std::from_chars(q, e, v, std::chars_format::general);   // correctly rounded, like Python's float()
    if (r.ec == std::errc::result_out_of_range) {
        // float() returns inf / 0.0 / a subnormal here; strtod agrees (rare path: "1e999", "1e-320")
        v = strtod(std::string(q, e).c_str(), nullptr);
    } else if (r.ec != std::errc() || r.ptr != e) {
        return TOK_BAD;
    }
    *out = neg ? -v : v;
    return TOK_OK;
}

// int(token): [+-]? digits
Tok read_int(const char *p, const char *e, long long *out, bool *out_of_range)
{
    const char *q = p;
    bool neg = false;
    if (q < e && (*q == '+' || *q == '-')) neg = (*q++ == '-');
    if (q == e) return exotic(p, e) ? TOK_REFUSE : TOK_BAD;
    long long v = 0;
    for (const char *s = q; s < e; ++s) {
        if (!is_digit(*s)) return exotic(p, e) ? TOK_REFUSE : TOK_BAD;
        if (v < (1LL << 40)) v = v * 10 + (*s - '0');
        else *out_of_range = true;
    }
    *out = neg ? -v : v;
    return TOK_OK;
}

struct Span {
    const char *b, *e;
};

void split_ws(const char *p, const char *e, std::vector<Span> &out)
{
    out.clear();
    while (p < e) {
        while (p < e && is_space((unsigned char)*p)) ++p;
        if (p == e) break;
        const char *s = p;
        while (p < e && !is_space((unsigned char)*p)) ++p;
        out.push_back({s, p});
    }
}

}  // namespace

struct crb_obj {
    std::vector<float> v, vt, vn;
    std::vector<int32_t> tv, tvt, tvn;
    std::vector<std::string> mtllibs;
    int64_t n_vt = 0;
    int vt_width = 0;          // -1: rows of different lengths
    bool has_tvt = true, has_tvn = true;
    int64_t good_lines = 0, bad_lines = 0, first_bad_line = 0;
};

namespace {

// One corner "v[/vt[/vn]]" -> the three fields of (corner + '//').split('/')[:3]
inline void corner_fields(Span c, Span f[3])
{
    const char *p = c.b;
    for (int k = 0; k < 3; ++k) {
        const char *s = p;
        while (p < c.e && *p != '/') ++p;
        f[k] = {s, p};
        if (p < c.e) ++p;   // skip the slash; at the end the appended "//" supplies empty fields
    }
}

// model.py:275-279
inline long long fix_index(long long i) { return i > 0 ? i - 1 : i; }

// Returns TOK_OK (line consumed), TOK_BAD (Python raises -> line skipped), TOK_REFUSE.
Tok read_face(const std::vector<Span> &comp, crb_obj *o, bool *range_error, std::vector<long long> &scratch)
{
    int n_tri = (int)comp.size() - 2;
    if (n_tri < 1) return TOK_OK;   // `range(len(comp) - 2)` is empty: nothing happens
    // scratch: per triangle 9 indices + 2 flags
    scratch.assign((size_t)n_tri * 11, 0);
    bool any_no_vt = false, any_no_vn = false;
    for (int t = 0; t < n_tri; ++t) {
        const Span cs[3] = {comp[0], comp[1 + t], comp[2 + t]};
        bool vt_ok = true, vn_ok = true;
        long long *row = &scratch[(size_t)t * 11];
        for (int c = 0; c < 3; ++c) {
            Span f[3];
            corner_fields(cs[c], f);
            long long i;
            Tok r = read_int(f[0].b, f[0].e, &i, range_error);
            if (r != TOK_OK) return r;
            row[c] = fix_index(i);
            if (f[1].b == f[1].e) vt_ok = false;
            if (vt_ok) {   // model.py:300-303: parsed only while this triangle's vt list is alive
                r = read_int(f[1].b, f[1].e, &i, range_error);
                if (r != TOK_OK) return r;
                row[3 + c] = fix_index(i);
            }
            if (f[2].b == f[2].e) vn_ok = false;
            if (vn_ok) {
                r = read_int(f[2].b, f[2].e, &i, range_error);
                if (r != TOK_OK) return r;
                row[6 + c] = fix_index(i);
            }
        }
        row[9] = vt_ok;
        row[10] = vn_ok;
        any_no_vt |= !vt_ok;
        any_no_vn |= !vn_ok;
    }
    // model.py:44-56
    auto fits = [&](long long x) {
        if (x < INT32_MIN || x > INT32_MAX) *range_error = true;
        return (int32_t)x;
    };
    for (int t = 0; t < n_tri; ++t)
        for (int c = 0; c < 3; ++c) o->tv.push_back(fits(scratch[(size_t)t * 11 + c]));
    if (any_no_vt) o->has_tvt = false;
    if (o->has_tvt)
        for (int t = 0; t < n_tri; ++t)
            for (int c = 0; c < 3; ++c) o->tvt.push_back(fits(scratch[(size_t)t * 11 + 3 + c]));
    if (any_no_vn) o->has_tvn = false;
    if (o->has_tvn)
        for (int t = 0; t < n_tri; ++t)
            for (int c = 0; c < 3; ++c) o->tvn.push_back(fits(scratch[(size_t)t * 11 + 6 + c]));
    return TOK_OK;
}

}  // namespace

extern "C" int crb_obj_parse(const char *text, size_t bytes, crb_obj **out)
{
    if (!out || (!text && bytes)) return failf(CRB_ERR_INVALID, "crb_obj_parse: NULL argument");
    *out = nullptr;
    crb_obj *o = new crb_obj();
    std::vector<Span> tok;
    std::vector<double> vals;
    std::vector<long long> scratch;
    bool range_error = false;
    const char *p = text, *end = text + bytes;
    int64_t line_no = 0;
    while (p < end) {
        // universal newlines: "\n", "\r\n", "\r"
        const char *ls = p;
        while (p < end && *p != '\n' && *p != '\r') ++p;
        const char *le = p;
        if (p < end) p += (*p == '\r' && p + 1 < end && p[1] == '\n') ? 2 : 1;
        ++line_no;
        if (ls == le || *ls == '#') continue;
        const char *sp = (const char *)memchr(ls, ' ', (size_t)(le - ls));
        if (!sp) continue;   // `line.split(' ', 1)` gives one part
        size_t clen = (size_t)(sp - ls);
        const char *data = sp + 1;
        auto cmd = [&](const char *w) { return clen == strlen(w) && memcmp(ls, w, clen) == 0; };
        Tok r = TOK_OK;
        if (cmd("v") || cmd("vt") || cmd("vn")) {
            split_ws(data, le, tok);
            vals.clear();
            for (const Span &t : tok) {
                double d;
                r = read_float(t.b, t.e, &d);
                if (r != TOK_OK) break;
                vals.push_back(d);
            }
            if (r == TOK_OK) {
                if (clen == 1) {   // v: model.py:258-261
                    if (vals.size() >= 3)
                        for (int k = 0; k < 3; ++k) o->v.push_back((float)vals[k]);
                    else r = TOK_BAD;   // the assert fails
                } else if (ls[1] == 't') {   // vt: model.py:264-265
                    if (o->n_vt == 0) o->vt_width = (int)vals.size();
                    else if (o->vt_width != (int)vals.size()) o->vt_width = -1;
                    ++o->n_vt;
                    if (o->vt_width >= 0)
                        for (double d : vals) o->vt.push_back((float)d);
                } else {   // vn: model.py:268-271
                    if (vals.size() == 3)
                        for (int k = 0; k < 3; ++k) o->vn.push_back((float)vals[k]);
                    else r = TOK_BAD;
                }
            }
        } else if (cmd("f")) {
            split_ws(data, le, tok);
            r = read_face(tok, o, &range_error, scratch);
        } else if (cmd("mtllib")) {
            o->mtllibs.emplace_back(data, (size_t)(le - data));
        }
        if (r == TOK_BAD) {   // model.py:71-74: the exception is swallowed (silent) and line_index not advanced
            if (o->bad_lines++ == 0) o->first_bad_line = o->good_lines + 1;
        } else {
            ++o->good_lines;
        }
        if (r == TOK_REFUSE) {
            delete o;
            return failf(CRB_ERR_SYNTAX, "line %lld: a numeric token with '_' or non-ASCII characters is not supported",
                         (long long)line_no);
        }
    }
    if (range_error) {
        delete o;
        return failf(CRB_ERR_RANGE, "Python integer out of bounds for int32");
    }
    *out = o;
    return CRB_OK;
}

extern "C" void crb_obj_free(crb_obj *o) { delete o; }

extern "C" int crb_obj_counts(const crb_obj *o, int64_t counts[10])
{
    if (!o || !counts) return failf(CRB_ERR_INVALID, "crb_obj_counts: NULL argument");
    counts[0] = (int64_t)o->v.size() / 3;
    counts[1] = o->n_vt;
    counts[2] = o->vt_width;
    counts[3] = (int64_t)o->vn.size() / 3;
    counts[4] = (int64_t)o->tv.size() / 3;
    counts[5] = o->has_tvt;
    counts[6] = o->has_tvn;
    counts[7] = (int64_t)o->mtllibs.size();
    counts[8] = o->bad_lines;
    counts[9] = o->first_bad_line;
    return CRB_OK;
}

extern "C" int crb_obj_copy(const crb_obj *o, float *v, float *vt, float *vn, int32_t *tri_v, int32_t *tri_vt,
                            int32_t *tri_vn)
{
    if (!o) return failf(CRB_ERR_INVALID, "crb_obj_copy: NULL object");
    if (v) memcpy(v, o->v.data(), o->v.size() * sizeof(float));
    if (vt) {
        if (o->vt_width < 0) return failf(CRB_ERR_INVALID, "texture coordinates of different lengths");
        memcpy(vt, o->vt.data(), o->vt.size() * sizeof(float));
    }
    if (vn) memcpy(vn, o->vn.data(), o->vn.size() * sizeof(float));
    if (tri_v) memcpy(tri_v, o->tv.data(), o->tv.size() * sizeof(int32_t));
    if (tri_vt) {
        if (!o->has_tvt) return failf(CRB_ERR_INVALID, "the faces carry no texture indices");
        memcpy(tri_vt, o->tvt.data(), o->tvt.size() * sizeof(int32_t));
    }
    if (tri_vn) {
        if (!o->has_tvn) return failf(CRB_ERR_INVALID, "the faces carry no normal indices");
        memcpy(tri_vn, o->tvn.data(), o->tvn.size() * sizeof(int32_t));
    }
    return CRB_OK;
}

extern "C" int crb_obj_mtllib(const crb_obj *o, int k, const char **data, size_t *len)
{
    if (!o || !data || !len || k < 0 || (size_t)k >= o->mtllibs.size())
        return failf(CRB_ERR_INVALID, "crb_obj_mtllib: bad argument");
    *data = o->mtllibs[(size_t)k].data();
    *len = o->mtllibs[(size_t)k].size();
    return CRB_OK;
}

// =============================================================================================================
// device side
// =============================================================================================================
namespace {

constexpr unsigned FULL = 0xffffffffu;

// np.dot / x.dot(x) of float32 3-vectors: float32 products, double accumulator, one narrowing (cblas_sdot's
// short-vector loop).
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    float p0 = __fmul_rn(ax, bx), p1 = __fmul_rn(ay, by), p2 = __fmul_rn(az, bz);
    double acc = 0.0;
    acc = __dadd_rn(acc, (double)p0);
    acc = __dadd_rn(acc, (double)p1);
    acc = __dadd_rn(acc, (double)p2);
    return __double2float_rn(acc);
}

// model.py:190-194
__device__ __forceinline__ void normalize3(float &x, float &y, float &z)
{
    float nrm = __fsqrt_rn(dot3(x, y, z, x, y, z));
    if (nrm == 0.0f) return;
    x = __fdiv_rn(x, nrm);
    y = __fdiv_rn(y, nrm);
    z = __fdiv_rn(z, nrm);
}

// Face normals (model.py:196-201) and the vertices' valences.
__global__ void __launch_bounds__(256) k_face_normals(const float *__restrict__ vert, const int32_t *__restrict__ tri,
                                                      int64_t T, float4 *__restrict__ faceN, int *__restrict__ cnt)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
    float t0[3], t1[3], t2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        t0[k] = vert[3 * (int64_t)i0 + k];
        t1[k] = vert[3 * (int64_t)i1 + k];
        t2[k] = vert[3 * (int64_t)i2 + k];
    }
    float a[3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = __fsub_rn(t1[k], t0[k]);
        b[k] = __fsub_rn(t1[k], t2[k]);
    }
    float nx = -__fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
    float ny = -__fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
    float nz = -__fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
    normalize3(nx, ny, nz);
    faceN[t] = make_float4(nx, ny, nz, 0.0f);
    atomicAdd(&cnt[i0], 1);
    atomicAdd(&cnt[i1], 1);
    atomicAdd(&cnt[i2], 1);
}

// Exclusive scan of cnt[0..n) in three small kernels: block-local scan + block totals, scan of the totals, add-back.
constexpr int SCAN_THREADS = 512, SCAN_PER_THREAD = 4, SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ int block_exclusive_scan(int x, int *total)
{
    __shared__ int warp_sums[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < nw ? warp_sums[lane] : 0, si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(FULL, si, d);
            if (lane >= d) si += y;
        }
        warp_sums[lane] = si - s;   // exclusive
        if (lane == 31) *total = si;
    }
    __syncthreads();
    int r = warp_sums[w] + inc - x;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_local(const int *__restrict__ cnt, int64_t n, int *__restrict__ off,
                                                             int *__restrict__ sums)
{
    __shared__ int total;
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        v[k] = base + k < n ? cnt[base + k] : 0;
        s += v[k];
    }
    int ex = block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        if (base + k < n) off[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(int *__restrict__ sums, int nblocks, int *__restrict__ grand)
{
    __shared__ int total;
    int carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += SCAN_THREADS) {
        int i = b0 + threadIdx.x;
        int x = i < nblocks ? sums[i] : 0;
        int ex = block_exclusive_scan(x, &total);
        if (i < nblocks) sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int *__restrict__ off, int *__restrict__ cursor, int64_t n,
                                                           const int *__restrict__ sums, const int *__restrict__ grand)
{
    int add = sums[blockIdx.x];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_PER_THREAD;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k)
        if (base + k < n) {
            int o = off[base + k] + add;
            off[base + k] = o;
            cursor[base + k] = o;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) off[n] = *grand;
}

// Incidence lists: entry e = 3*triangle + corner lands somewhere in its vertex's segment (order fixed later).
__global__ void __launch_bounds__(256) k_fill_incidence(const int32_t *__restrict__ tri, int64_t n_entries,
                                                        int *__restrict__ cursor, int *__restrict__ inc)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    int slot = atomicAdd(&cursor[tri[e]], 1);
    inc[slot] = (int)e;
}

// One warp per vertex (model.py:176-188).  The vertex's entries are first put in file order (rank sort: entries are
// distinct), then walked in that order: a face normal is kept unless a kept one has dot >= 1 with it.  The first 32
// kept normals live in registers (lane j holds the j-th), later ones in the `kept` scratch.  The mean adds the kept
// normals in order, starting from +0.0 like np.add.reduce.
// Vertices met by more than HEAVY_VALENCE corners are handed to k_vertex_normals_heavy (one CTA each).
constexpr int HEAVY_VALENCE = 512;

__device__ __forceinline__ void finish_vertex(float sx, float sy, float sz, int m, int invert, float *out)
{
    if (m > 0) {
        double dm = (double)m;
        sx = __double2float_rn(__ddiv_rn((double)sx, dm));
        sy = __double2float_rn(__ddiv_rn((double)sy, dm));
        sz = __double2float_rn(__ddiv_rn((double)sz, dm));
        normalize3(sx, sy, sz);
    }
    if (invert) {   // model.py:168-169  `self._normals *= -1`
        sx = __fmul_rn(sx, -1.0f);
        sy = __fmul_rn(sy, -1.0f);
        sz = __fmul_rn(sz, -1.0f);
    }
    out[0] = sx;
    out[1] = sy;
    out[2] = sz;
}

__global__ void __launch_bounds__(256) k_vertex_normals(const float4 *__restrict__ faceN, const int *__restrict__ off,
                                                        const int *__restrict__ inc, int *__restrict__ sorted,
                                                        float *__restrict__ kept, int64_t V, int invert,
                                                        float *__restrict__ out, int *__restrict__ heavy_count,
                                                        int *__restrict__ heavy_list)
{
    int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (v >= V) return;
    int base = off[v], k = off[v + 1] - base;
    if (k > HEAVY_VALENCE) {
        if (lane == 0) heavy_list[atomicAdd(heavy_count, 1)] = (int)v;
        return;
    }
    for (int i = lane; i < k; i += 32) {
        int e = inc[base + i], r = 0;
        for (int j = 0; j < k; ++j) r += inc[base + j] < e;
        __stcg(&sorted[base + r], e);
    }
    __syncwarp();
    float kx = 0.f, ky = 0.f, kz = 0.f;      // this lane's kept normal
    float sx = 0.f, sy = 0.f, sz = 0.f;      // ordered sum (every lane keeps the same copy)
    int m = 0;
    for (int i = 0; i < k; ++i) {
        int e = __ldcg(&sorted[base + i]);
        float4 n = faceN[e / 3];
        bool dup = false;
        if (lane < m) dup = dot3(kx, ky, kz, n.x, n.y, n.z) >= 1.0f;
        for (int j = 32 + lane; j < m; j += 32) {
            const float *q = kept + 3 * ((int64_t)base + j);
            dup |= dot3(__ldcg(q), __ldcg(q + 1), __ldcg(q + 2), n.x, n.y, n.z) >= 1.0f;
        }
        if (__any_sync(FULL, dup)) continue;
        if (m < 32) {
            if (lane == m) kx = n.x, ky = n.y, kz = n.z;
        } else if (lane == 0) {
            float *q = kept + 3 * ((int64_t)base + m);
            __stcg(q, n.x);
            __stcg(q + 1, n.y);
            __stcg(q + 2, n.z);
        }
        ++m;
        sx = __fadd_rn(sx, n.x);
        sy = __fadd_rn(sy, n.y);
        sz = __fadd_rn(sz, n.z);
        __syncwarp();
    }
    if (lane == 0) finish_vertex(sx, sy, sz, m, invert, out + 3 * v);
}

// One CTA per heavy vertex (a fan centre, a sphere pole): the same walk as above with the work of each step spread over
// 1024 threads.  Entries are put in file order by a rank sort through shared-memory tiles; the walk keeps the kept
// normals in shared memory (HEAVY_KEPT_SMEM of them, later ones in the `kept` scratch) and costs one block barrier per
// entry: the normal appended by the previous step is compared from registers, so its store need not be visible yet.
constexpr int HEAVY_THREADS = 1024, HEAVY_TILE = 4096, HEAVY_STAGE = 1024, HEAVY_KEPT_SMEM = 12288;
constexpr size_t HEAVY_SMEM_BYTES = sizeof(float) * 3 * (HEAVY_KEPT_SMEM + HEAVY_STAGE);

__global__ void __launch_bounds__(HEAVY_THREADS) k_vertex_normals_heavy(
    const float4 *__restrict__ faceN, const int *__restrict__ off, const int *__restrict__ inc, int *__restrict__ sorted,
    float *__restrict__ kept, int invert, float *__restrict__ out, const int *__restrict__ heavy_count,
    const int *__restrict__ heavy_list)
{
    extern __shared__ float smem[];
    float *skept = smem;                                   // [HEAVY_KEPT_SMEM][3]
    float *stage = smem + 3 * HEAVY_KEPT_SMEM;             // [HEAVY_STAGE][3]
    int *tile = reinterpret_cast<int *>(smem);             // rank-sort tiles alias the kept area (used before the walk)
    const int tid = threadIdx.x;
    const int n_heavy = *heavy_count;
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const int v = heavy_list[h];
        const int base = off[v], k = off[v + 1] - base;
        // ---- file order: rank of every entry among the vertex's entries
        for (int i0 = 0; i0 < k; i0 += HEAVY_THREADS) {
            int i = i0 + tid, e = i < k ? inc[base + i] : 0, r = 0;
            for (int j0 = 0; j0 < k; j0 += HEAVY_TILE) {
                int nt = min(HEAVY_TILE, k - j0);
                __syncthreads();
                for (int j = tid; j < nt; j += HEAVY_THREADS) tile[j] = inc[base + j0 + j];
                __syncthreads();
                if (i < k)
                    for (int j = 0; j < nt; ++j) r += tile[j] < e;
            }
            if (i < k) sorted[base + r] = e;
        }
        __syncthreads();
        // ---- the walk
        float sx = 0.f, sy = 0.f, sz = 0.f, lx = 0.f, ly = 0.f, lz = 0.f;
        int m = 0;
        bool pending = false;   // kept[m-1] was stored by thread 0 after the last barrier: compare `l` instead
        for (int i0 = 0; i0 < k; i0 += HEAVY_STAGE) {
            int ns = min(HEAVY_STAGE, k - i0);
            __syncthreads();
            if (tid < ns) {
                float4 n = faceN[sorted[base + i0 + tid] / 3];
                stage[3 * tid] = n.x;
                stage[3 * tid + 1] = n.y;
                stage[3 * tid + 2] = n.z;
            }
            __syncthreads();
            for (int i = 0; i < ns; ++i) {
                const float nx = stage[3 * i], ny = stage[3 * i + 1], nz = stage[3 * i + 2];
                const int lim = m - (pending ? 1 : 0);
                bool dup = false;
                for (int j = tid; j < lim; j += HEAVY_THREADS) {
                    float qx, qy, qz;
                    if (j < HEAVY_KEPT_SMEM) {
                        qx = skept[3 * j], qy = skept[3 * j + 1], qz = skept[3 * j + 2];
                    } else {
                        const float *q = kept + 3 * ((int64_t)base + j);
                        qx = __ldcg(q), qy = __ldcg(q + 1), qz = __ldcg(q + 2);
                    }
                    dup |= dot3(qx, qy, qz, nx, ny, nz) >= 1.0f;
                }
                if (pending && tid == 0) dup |= dot3(lx, ly, lz, nx, ny, nz) >= 1.0f;
                const int any = __syncthreads_or(dup);
                pending = false;
                if (any) continue;
                if (tid == 0) {
                    if (m < HEAVY_KEPT_SMEM) {
                        skept[3 * m] = nx, skept[3 * m + 1] = ny, skept[3 * m + 2] = nz;
                    } else {
                        float *q = kept + 3 * ((int64_t)base + m);
                        __stcg(q, nx);
                        __stcg(q + 1, ny);
                        __stcg(q + 2, nz);
                    }
                }
                lx = nx, ly = ny, lz = nz;
                pending = true;
                ++m;
                sx = __fadd_rn(sx, nx);
                sy = __fadd_rn(sy, ny);
                sz = __fadd_rn(sz, nz);
            }
        }
        if (tid == 0) finish_vertex(sx, sy, sz, m, invert, out + 3 * (int64_t)v);
        __syncthreads();
    }
}

// astype('int32') of a float32 on x86-64 (cvttps2dq): INT32_MIN for NaN and out-of-range values.
__device__ __forceinline__ int f32_to_i32_x86(float x)
{
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return INT_MIN;
    return (int)x;
}

// model.py:147-150
__global__ void __launch_bounds__(256) k_vertex_colors(const float *__restrict__ vt, int64_t n, int width,
                                                       const uint8_t *__restrict__ tex, int h, int w,
                                                       float *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float fy = __fmul_rn(__fsub_rn(1.0f, vt[width * i + 1]), (float)h);
    float fx = __fmul_rn(vt[width * i], (float)w);
    int y = min(max(f32_to_i32_x86(fy), 0), h - 1), x = min(max(f32_to_i32_x86(fx), 0), w - 1);
    const uint8_t *px = tex + 3 * ((int64_t)y * w + x);
    out[3 * i] = (float)px[0];
    out[3 * i + 1] = (float)px[1];
    out[3 * i + 2] = (float)px[2];
}

// model.py:151,158,172: one thread per output float, coalesced stores.
__global__ void __launch_bounds__(256) k_gather_by_triangles(const float *__restrict__ attr, const int32_t *__restrict__ tri,
                                                             int64_t n_floats, float *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_floats) return;
    int64_t e = i / 3;
    int k = (int)(i - 3 * e);
    out[i] = attr[3 * (int64_t)tri[e] + k];
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct NormalsWs {
    float4 *faceN;
    int *cnt, *off, *cursor, *inc, *sorted, *sums, *grand, *heavy_count, *heavy_list;
    float *kept;
    size_t bytes;
};

NormalsWs carve(char *p, int64_t V, int64_t T)
{
    NormalsWs w;
    size_t o = 0;
    auto take = [&](size_t n) {
        size_t at = o;
        o = align_up(o + n, 256);
        return p + at;
    };
    int64_t nb = (V + SCAN_TILE - 1) / SCAN_TILE;
    w.faceN = (float4 *)take(sizeof(float4) * (size_t)(T ? T : 1));
    w.cnt = (int *)take(sizeof(int) * (size_t)(V + 1));
    w.off = (int *)take(sizeof(int) * (size_t)(V + 1));
    w.cursor = (int *)take(sizeof(int) * (size_t)(V + 1));
    w.inc = (int *)take(sizeof(int) * (size_t)(3 * T + 1));
    w.sorted = (int *)take(sizeof(int) * (size_t)(3 * T + 1));
    w.kept = (float *)take(sizeof(float) * (size_t)(9 * T + 3));
    w.sums = (int *)take(sizeof(int) * (size_t)(nb + 1));
    w.grand = (int *)take(sizeof(int));
    w.heavy_count = (int *)take(sizeof(int));
    w.heavy_list = (int *)take(sizeof(int) * (size_t)(3 * T / HEAVY_VALENCE + 1));
    w.bytes = o;
    return w;
}

inline unsigned blocks_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

}  // namespace

extern "C" size_t crb_model_normals_workspace_bytes(int64_t V, int64_t T)
{
    if (V < 0 || T < 0) return 0;
    return carve(nullptr, V, T).bytes;
}

extern "C" int crb_model_vertex_normals(const float *vertices, int64_t V, const int32_t *tri, int64_t T, int invert,
                                        float *normals_out, void *workspace, size_t workspace_bytes, void *stream)
{
    if (V < 0 || T < 0 || (V > 0 && (!vertices || !normals_out)) || (T > 0 && !tri))
        return failf(CRB_ERR_INVALID, "crb_model_vertex_normals: bad argument");
    if (3 * T > (int64_t)INT32_MAX - 8 || V > (int64_t)INT32_MAX / 64)
        return failf(CRB_ERR_INVALID, "crb_model_vertex_normals: mesh too large (3T must fit in int32)");
    if (V == 0) return CRB_OK;
    NormalsWs w = carve((char *)workspace, V, T);
    if (!workspace || workspace_bytes < w.bytes)
        return failf(CRB_ERR_STATE, "crb_model_vertex_normals: workspace of %zu bytes needed, %zu given", w.bytes,
                     workspace_bytes);
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemsetAsync(w.cnt, 0, sizeof(int) * (size_t)(V + 1), s));
    int launches = 0;
    if (T > 0) {
        k_face_normals<<<blocks_for(T, 256), 256, 0, s>>>(vertices, tri, T, w.faceN, w.cnt);
        ++launches;
    }
    unsigned nb = blocks_for(V, SCAN_TILE);
    k_scan_local<<<nb, SCAN_THREADS, 0, s>>>(w.cnt, V, w.off, w.sums);
    k_scan_sums<<<1, SCAN_THREADS, 0, s>>>(w.sums, (int)nb, w.grand);
    k_scan_add<<<nb, SCAN_THREADS, 0, s>>>(w.off, w.cursor, V, w.sums, w.grand);
    launches += 3;
    if (T > 0) {
        k_fill_incidence<<<blocks_for(3 * T, 256), 256, 0, s>>>(tri, 3 * T, w.cursor, w.inc);
        ++launches;
    }
    CU(cudaMemsetAsync(w.heavy_count, 0, sizeof(int), s));
    k_vertex_normals<<<blocks_for(V * 32, 256), 256, 0, s>>>(w.faceN, w.off, w.inc, w.sorted, w.kept, V, invert ? 1 : 0,
                                                            normals_out, w.heavy_count, w.heavy_list);
    ++launches;
    if (3 * T > HEAVY_VALENCE) {   // otherwise no vertex can be heavy
        CU(cudaFuncSetAttribute(k_vertex_normals_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)HEAVY_SMEM_BYTES));   // per device, cheap
        unsigned grid = (unsigned)std::min<int64_t>(3 * T / HEAVY_VALENCE, 148);
        k_vertex_normals_heavy<<<grid, HEAVY_THREADS, HEAVY_SMEM_BYTES, s>>>(w.faceN, w.off, w.inc, w.sorted, w.kept,
                                                                             invert ? 1 : 0, normals_out, w.heavy_count,
                                                                             w.heavy_list);
        ++launches;
    }
    CU(cudaGetLastError());
    g_launches += launches;
    return CRB_OK;
}

extern "C" int crb_model_vertex_colors(const float *vt, int64_t n, int width, const uint8_t *texture, int tex_h,
                                       int tex_w, float *colors_out, void *stream)
{
    if (n < 0 || width < 2 || tex_h <= 0 || tex_w <= 0 || !texture || (n > 0 && (!vt || !colors_out)))
        return failf(CRB_ERR_INVALID, "crb_model_vertex_colors: bad argument");
    if (n == 0) return CRB_OK;
    k_vertex_colors<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(vt, n, width, texture, tex_h, tex_w, colors_out);
    CU(cudaGetLastError());
    ++g_launches;
    return CRB_OK;
}

extern "C" int crb_model_gather(const float *attr, const int32_t *tri, int64_t T, float *out, void *stream)
{
    if (T < 0 || (T > 0 && (!attr || !tri || !out))) return failf(CRB_ERR_INVALID, "crb_model_gather: bad argument");
    if (T == 0) return CRB_OK;
    k_gather_by_triangles<<<blocks_for(9 * T, 256), 256, 0, (cudaStream_t)stream>>>(attr, tri, 9 * T, out);
    CU(cudaGetLastError());
    ++g_launches;
    return CRB_OK;
}

extern "C" int64_t crb_model_launch_count(void) { return g_launches.load(); }
