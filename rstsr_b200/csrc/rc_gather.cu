// rc_gather.cu -- index-driven data movement: index_select, pack_tri, unpack_tri.
//
//   rc_index_select  c[.., i, ..] = a[.., indices[i], ..]   DeviceIndexSelectAPI::index_select
//                    (rstsr-core/src/device_cpu_serial/adv_indexing.rs:3-20,
//                     rstsr-native-impl/src/cpu_serial/adv_indexing.rs:3-90)
//   rc_pack_tri      packed[.., p] = full[.., i, j]         OpPackTriAPI::pack_tri
//   rc_unpack_tri    full[.., i, j] = +-packed[.., p]       OpUnpackTriAPI::unpack_tri
//                    (rstsr-core/src/device_cpu_serial/operators/op_tri.rs:4-53,
//                     rstsr-native-impl/src/cpu_serial/op_tri.rs:7-522)
// All three are pure byte movement: HBM-bound, 2 x itemsize bytes per moved element.  Elements move as raw
// 1/2/4/8/16/32-byte words; only the antisymmetric unpack (negation) is typed (f32 / f64, as ComplexFloat is in
// the reference).
#include <cmath>

#include "rc_canon.hpp"
#include "rc_device.hpp"
#include "rc_kernel_common.cuh"
#include "rc_layout.hpp"

namespace rc {
namespace {

constexpr int GT_BLOCK = 256;
constexpr int GT_ITEMS = 4;  // independent moves in flight per thread

// the dims that are not indexed / not part of the triangle: joint canonical form of (output, input)
struct RestDesc {
    int nd;
    int big;  // some extent or the rest size needs 64-bit division
    int64_t shape[KMAXD], s_out[KMAXD], s_in[KMAXD];
    FastDiv dv[KMAXD];
    int64_t n_rest;
};

__device__ __forceinline__ void rest_offsets(const RestDesc &d, int64_t r, int64_t &o_out, int64_t &o_in) {
    o_out = 0;
    o_in = 0;
    if (d.big) {
        for (int k = 0; k < d.nd; ++k) {
            const int64_t q = r / d.shape[k], m = r - q * d.shape[k];
            o_out += m * d.s_out[k];
            o_in += m * d.s_in[k];
            r = q;
        }
        return;
    }
    uint32_t t = (uint32_t)r;
#pragma unroll 1
    for (int k = 0; k < d.nd; ++k) {
        uint32_t q, m;
        if (k + 1 < d.nd) d.dv[k].divmod(t, q, m); else { q = 0; m = t; }
        o_out += (int64_t)m * d.s_out[k];
        o_in += (int64_t)m * d.s_in[k];
        t = q;
    }
}

// t -> (t / div, t % div): 32-bit multiply-shift when the launch has fewer than 2^31 items
__device__ __forceinline__ void split_index(int64_t t, int64_t div, const FastDiv &fd, int big, int64_t &q, int64_t &m) {
    if (big) { q = t / div; m = t - q * div; return; }
    uint32_t q32, m32;
    fd.divmod((uint32_t)t, q32, m32);
    q = q32;
    m = m32;
}

// largest i with i (i + 1) / 2 <= p
__device__ __forceinline__ int64_t tri_row(int64_t p) {
    int64_t i = (int64_t)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while (i * (i + 1) / 2 > p) --i;
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    return i;
}

// ---------------- index_select ----------------
struct SelDesc {
    RestDesc rest;
    int64_t n_idx, sc_axis, sa_axis;
    int64_t n_src;    // extent of the indexed axis in the input
    int idx_fastest;  // threads walk the indexed axis first (it is the output's contiguous axis)
    int64_t total;
    FastDiv split;    // by n_idx (idx_fastest) or n_rest
};

template <class U>
__global__ void __launch_bounds__(GT_BLOCK) index_select_kernel(const __grid_constant__ SelDesc d, U *__restrict__ c,
                                                                const U *__restrict__ a,
                                                                const int64_t *__restrict__ idx) {
    const int64_t t0 = (int64_t)blockIdx.x * (GT_BLOCK * GT_ITEMS) + threadIdx.x;
    int64_t oc[GT_ITEMS], oa[GT_ITEMS];
    bool ok[GT_ITEMS];
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u) {
        const int64_t t = t0 + (int64_t)u * GT_BLOCK;
        ok[u] = t < d.total;
        int64_t i, r;
        if (d.idx_fastest) split_index(t, d.n_idx, d.split, d.rest.big, r, i);
        else split_index(t, d.rest.n_rest, d.split, d.rest.big, i, r);
        int64_t ro = 0, ri = 0;
        if (ok[u]) {
            rest_offsets(d.rest, r, ro, ri);
            oc[u] = ro + i * d.sc_axis;
            oa[u] = ri + idx[i] * d.sa_axis;
        }
    }
    U v[GT_ITEMS];
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u)
        if (ok[u]) v[u] = a[oa[u]];
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u)
        if (ok[u]) c[oc[u]] = v[u];
}

// Gather ALONG the contiguous axis (take(indices, -1) of a row-major tensor): random 8-byte reads fetch a 32-byte
// sector each (ncu: 31 % DRAM, 45 % L2 hit, long-scoreboard bound).  When a source row fits in shared memory and most
// of it is wanted anyway, one CTA stages the row with coalesced loads and gathers from shared memory instead.
template <class U>
__global__ void __launch_bounds__(GT_BLOCK) index_select_smem_kernel(const __grid_constant__ SelDesc d, U *__restrict__ c,
                                                                     const U *__restrict__ a,
                                                                     const int64_t *__restrict__ idx) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    U *row = reinterpret_cast<U *>(sel_smem);
    int64_t ro, ri;
    rest_offsets(d.rest, blockIdx.x, ro, ri);
    const U *src = a + ri;
    U *dst = c + ro;
    const int n_src = (int)d.n_src, n_idx = (int)d.n_idx;
    int t = threadIdx.x;
    for (; t + (GT_ITEMS - 1) * GT_BLOCK < n_src; t += GT_ITEMS * GT_BLOCK) {
        U v[GT_ITEMS];
#pragma unroll
        for (int u = 0; u < GT_ITEMS; ++u) v[u] = src[t + u * GT_BLOCK];
#pragma unroll
        for (int u = 0; u < GT_ITEMS; ++u) row[t + u * GT_BLOCK] = v[u];
    }
    for (; t < n_src; t += GT_BLOCK) row[t] = src[t];
    __syncthreads();
    t = threadIdx.x;
    for (; t + (GT_ITEMS - 1) * GT_BLOCK < n_idx; t += GT_ITEMS * GT_BLOCK) {
        int64_t k[GT_ITEMS];
#pragma unroll
        for (int u = 0; u < GT_ITEMS; ++u) k[u] = idx[t + u * GT_BLOCK];
#pragma unroll
        for (int u = 0; u < GT_ITEMS; ++u) dst[(int64_t)(t + u * GT_BLOCK) * d.sc_axis] = row[k[u]];
    }
    for (; t < n_idx; t += GT_BLOCK) dst[(int64_t)t * d.sc_axis] = row[idx[t]];
}

// The same gather when a row does NOT fit in shared memory but the indices are locally clustered (monotone selections:
// every other column, a mask, a few columns dropped): the output is cut into windows of `win_w` consecutive indices, the
// host records the source span [lo, lo + span) each window touches, and a CTA stages just that span (coalesced) before it
// gathers from shared memory and writes its window (coalesced).  Random indices over a long row keep the direct kernel:
// every 8-byte read is a 32-byte L2 sector either way, and sorting would cost a second pass.
template <class U>
__global__ void __launch_bounds__(GT_BLOCK) index_select_window_kernel(const __grid_constant__ SelDesc d, U *__restrict__ c,
                                                                       const U *__restrict__ a,
                                                                       const int64_t *__restrict__ idx,
                                                                       const int64_t *__restrict__ win, int win_w,
                                                                       FastDiv div_chunks) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    U *buf = reinterpret_cast<U *>(sel_smem);
    uint32_t r, chunk;
    div_chunks.divmod(blockIdx.x, r, chunk);
    int64_t ro, ri;
    rest_offsets(d.rest, r, ro, ri);
    const int64_t lo = win[2 * chunk];
    const int span = (int)win[2 * chunk + 1];
    const U *src = a + ri + lo;
    int t = threadIdx.x;
    for (; t + (GT_ITEMS - 1) * GT_BLOCK < span; t += GT_ITEMS * GT_BLOCK) {
        U v[GT_ITEMS];
#pragma unroll
        for (int u = 0; u < GT_ITEMS; ++u) v[u] = src[t + u * GT_BLOCK];
#pragma unroll
        for (int u = 0; u < GT_ITEMS; ++u) buf[t + u * GT_BLOCK] = v[u];
    }
    for (; t < span; t += GT_BLOCK) buf[t] = src[t];
    __syncthreads();
    const int64_t i0 = (int64_t)chunk * win_w;
    const int n = (int)(d.n_idx - i0 < win_w ? d.n_idx - i0 : win_w);
    U *dst = c + ro + i0 * d.sc_axis;
    const int64_t *ix = idx + i0;
    for (t = threadIdx.x; t < n; t += GT_BLOCK) dst[(int64_t)t * d.sc_axis] = buf[ix[t] - lo];
}

// ---------------- pack_tri ----------------
struct TriMoveDesc {
    RestDesc rest;       // s_out: strides of the OUTPUT operand, s_in: of the input
    int64_t n, n_tp;
    int64_t sp;          // stride of the packed axis
    int64_t si, sj;      // strides of the (i, j) axes of the full matrix
    int upper;           // packed triangle: 0 lower (j <= i), 1 upper (j >= i), both row by row
    int symm;            // rc_symm (unpack only)
    int tri_fastest;     // threads walk the packed axis first
    int64_t total;
    int64_t ntile, ntile_tri;  // unpack: tiles per side, tiles in the triangle
    FastDiv split;       // by n_tp (tri_fastest) or n_rest
    FastDiv split_n;     // by n (rest-fastest unpack)
};

template <class U>
__global__ void __launch_bounds__(GT_BLOCK) pack_tri_kernel(const __grid_constant__ TriMoveDesc d, U *__restrict__ packed,
                                                            const U *__restrict__ full) {
    const int64_t t0 = (int64_t)blockIdx.x * (GT_BLOCK * GT_ITEMS) + threadIdx.x;
    int64_t op[GT_ITEMS], of[GT_ITEMS];
    bool ok[GT_ITEMS];
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u) {
        const int64_t t = t0 + (int64_t)u * GT_BLOCK;
        ok[u] = t < d.total;
        if (!ok[u]) continue;
        int64_t p, r;
        if (d.tri_fastest) split_index(t, d.n_tp, d.split, d.rest.big, r, p);
        else split_index(t, d.rest.n_rest, d.split, d.rest.big, p, r);
        int64_t ro, ri;
        rest_offsets(d.rest, r, ro, ri);
        // upper triangle row by row == lower triangle of the point-mirrored matrix, walked backwards
        const int64_t q = d.upper ? d.n_tp - 1 - p : p;
        int64_t i = tri_row(q), j = q - i * (i + 1) / 2;
        if (d.upper) { i = d.n - 1 - i; j = d.n - 1 - j; }
        op[u] = ro + p * d.sp;
        of[u] = ri + i * d.si + j * d.sj;
    }
    U v[GT_ITEMS];
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u)
        if (ok[u]) v[u] = full[of[u]];
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u)
        if (ok[u]) packed[op[u]] = v[u];
}

// ---------------- tile kernels for the triangle ----------------
// One CTA per TT x TT tile of the STORED triangle (tiles of the other triangle are never visited) and per rest
// index; a packed row is a contiguous run along j, so both the packed and the full side move in 512-byte row
// segments.  All loads of a thread are issued before its first store.
// TT = 64: 4 rows per pass, 16 passes; TT = 32: 8 rows per pass, 4 passes (RC_TRI_TILE / RC_UNPACK_TILE select;
// measured with the interior fast path: pack 6.2 vs 5.6 TB/s, unpack 5.9 vs 5.4 TB/s in favour of 64)
inline int tri_tile() {
    static int v = [] { const char *e = getenv("RC_TRI_TILE"); int x = e ? atoi(e) : 64; return x == 32 ? 32 : 64; }();
    return v;
}
inline int unpack_tile() {
    static int v = [] { const char *e = getenv("RC_UNPACK_TILE"); int x = e ? atoi(e) : 64; return x == 32 ? 32 : 64; }();
    return v;
}

struct TileCoord { int64_t ti, tj, ro, ri; };

template <int TT>
__device__ __forceinline__ TileCoord tile_coord(const TriMoveDesc &d) {
    const int64_t blk = blockIdx.x;
    const int64_t r = blk / d.ntile_tri;
    const int64_t q = blk - r * d.ntile_tri;
    const int64_t x = tri_row(q), y = q - x * (x + 1) / 2;  // x >= y
    TileCoord t;
    t.ti = d.upper ? y : x;
    t.tj = d.upper ? x : y;
    rest_offsets(d.rest, r, t.ro, t.ri);
    return t;
}

__device__ __forceinline__ int64_t packed_index(const TriMoveDesc &d, int64_t i, int64_t j) {
    return d.upper ? i * d.n - i * (i - 1) / 2 + (j - i) : i * (i + 1) / 2 + j;
}

// Offsets (pre-multiplied by the packed stride) of the packed elements (i0 + k T, j), k = 0, 1, ..: the index is
// quadratic in i, so it advances by a first difference that itself advances by a constant -- two adds per element
// instead of 64-bit multiplies (ncu: the multiply version was issue-bound at 68-74 % issue-active, 47 % DRAM).
struct PackedWalk { int64_t q, dq, ddq; };
__device__ __forceinline__ PackedWalk packed_walk(const TriMoveDesc &d, int64_t i0, int64_t j, int64_t T) {
    PackedWalk w;
    if (d.upper) {
        w.q = i0 * d.n - i0 * (i0 - 1) / 2 + (j - i0);
        w.dq = T * d.n - T * i0 - T * (T - 1) / 2 - T;
        w.ddq = -T * T;
    } else {
        w.q = i0 * (i0 + 1) / 2 + j;
        w.dq = T * i0 + T * (T + 1) / 2;
        w.ddq = T * T;
    }
    w.q *= d.sp; w.dq *= d.sp; w.ddq *= d.sp;
    return w;
}

// tiles strictly off the diagonal and inside the matrix need no per-element predicate
template <int TT>
__device__ __forceinline__ bool interior_tile(const TriMoveDesc &d, const TileCoord &t) {
    return t.ti != t.tj && (t.ti + 1) * TT <= d.n && (t.tj + 1) * TT <= d.n;
}

// pack_tri when the packed axis is the output's fastest: packed[p(i, j)] = full[i, j]
template <class U, int TT>
__global__ void __launch_bounds__(GT_BLOCK) pack_tri_tile_kernel(const __grid_constant__ TriMoveDesc d,
                                                                 U *__restrict__ packed, const U *__restrict__ full) {
    constexpr int TROWS = GT_BLOCK / TT, TPASS = TT / TROWS;
    const TileCoord t = tile_coord<TT>(d);
    U *dst = packed + t.ro;
    const U *src = full + t.ri;
    const int tx = threadIdx.x % TT, ty = threadIdx.x / TT;
    const int64_t j = t.tj * TT + tx;
    U v[TPASS];
    if (interior_tile<TT>(d, t)) {
        const int64_t i0 = t.ti * TT + ty;
        const U *s = src + i0 * d.si + j * d.sj;
        const int64_t sstep = (int64_t)TROWS * d.si;
#pragma unroll
        for (int k = 0; k < TPASS; ++k) { v[k] = *s; s += sstep; }
        PackedWalk w = packed_walk(d, i0, j, TROWS);
#pragma unroll
        for (int k = 0; k < TPASS; ++k) { dst[w.q] = v[k]; w.q += w.dq; w.dq += w.ddq; }
        return;
    }
#pragma unroll
    for (int k = 0; k < TPASS; ++k) {
        const int64_t i = t.ti * TT + ty + k * TROWS;
        const bool in_tri = i < d.n && j < d.n && (d.upper ? j >= i : j <= i);
        if (in_tri) v[k] = src[i * d.si + j * d.sj];
    }
#pragma unroll
    for (int k = 0; k < TPASS; ++k) {
        const int64_t i = t.ti * TT + ty + k * TROWS;
        const bool in_tri = i < d.n && j < d.n && (d.upper ? j >= i : j <= i);
        if (in_tri) dst[packed_index(d, i, j) * d.sp] = v[k];
    }
}

// unpack_tri: the tile is read from the packed rows, written to full[i, j], staged in shared memory and written
// transposed to full[j, i] with the symmetry's sign -- every element of the full matrix is written at most once,
// by exactly one thread.
template <class T, int TT>
__global__ void __launch_bounds__(GT_BLOCK) unpack_tri_kernel(const __grid_constant__ TriMoveDesc d, T *__restrict__ full,
                                                              const T *__restrict__ packed) {
    constexpr int TROWS = GT_BLOCK / TT, TPASS = TT / TROWS;
    __shared__ T tile[TT][TT + 1];
    const TileCoord t = tile_coord<TT>(d);
    T *dst = full + t.ro;
    const T *src = packed + t.ri;
    const int tx = threadIdx.x % TT, ty = threadIdx.x / TT;
    const bool anti = d.symm == RC_SYMM_AY || d.symm == RC_SYMM_AH;
    const int64_t j = t.tj * TT + tx;
    if (interior_tile<TT>(d, t)) {
        const int64_t i0 = t.ti * TT + ty;
        const int64_t ostep = (int64_t)TROWS * d.si;
        PackedWalk w = packed_walk(d, i0, j, TROWS);
        T *o = dst + i0 * d.si + j * d.sj;
#pragma unroll
        for (int k = 0; k < TPASS; ++k) {
            const T v = src[w.q];
            *o = v;
            tile[ty + k * TROWS][tx] = v;
            o += ostep; w.q += w.dq; w.dq += w.ddq;
        }
        if (d.symm == RC_SYMM_N) return;
        __syncthreads();
        T *m = dst + (t.tj * TT + ty) * d.si + (t.ti * TT + tx) * d.sj;
#pragma unroll
        for (int k = 0; k < TPASS; ++k) {
            T v = tile[tx][ty + k * TROWS];
            if (anti) v = -v;
            *m = v;
            m += ostep;
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < TPASS; ++k) {
        const int li = ty + k * TROWS;
        const int64_t i = t.ti * TT + li;
        const bool in_tri = i < d.n && j < d.n && (d.upper ? j >= i : j <= i);
        T v{};
        if (in_tri) {
            v = src[packed_index(d, i, j) * d.sp];
            if (anti && i == j) v = T{};  // the stored diagonal is ignored: written as zero
            dst[i * d.si + j * d.sj] = v;
        }
        tile[li][tx] = v;
    }
    if (d.symm == RC_SYMM_N) return;  // the other triangle stays untouched (uninitialised in the reference)
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TPASS; ++k) {
        const int lj = ty + k * TROWS;
        // mirrored element full[j, i] of source (i, j): i along tx now, so the store walks full's last axis
        const int64_t i = t.ti * TT + tx, jj = t.tj * TT + lj;
        const bool in_tri = i < d.n && jj < d.n && (d.upper ? jj > i : jj < i);
        if (in_tri) {
            T w = tile[tx][lj];
            if (anti) w = -w;
            dst[jj * d.si + i * d.sj] = w;
        }
    }
}

// unpack_tri when the batch ("rest") axis is the memory-fastest one (e.g. a [.., n, n].f() tensor on a row-major
// handle, the layout of the reference's own rayon test): every (i, j) of the FULL matrix owns a contiguous run of
// batch elements, so one thread moves one 32-byte pack of that run from the packed element it mirrors.  (The tile
// kernel would touch one sector per element there: measured 0.9 TB/s.)
template <class T, int V>
__global__ void __launch_bounds__(GT_BLOCK) unpack_tri_rest_kernel(const __grid_constant__ TriMoveDesc d, T *__restrict__ full,
                                                                   const T *__restrict__ packed) {
    const bool anti = d.symm == RC_SYMM_AY || d.symm == RC_SYMM_AH;
    const int64_t t0 = (int64_t)blockIdx.x * (GT_BLOCK * GT_ITEMS) + threadIdx.x;
    int64_t of[GT_ITEMS], op[GT_ITEMS];
    int act[GT_ITEMS];  // 0: nothing to write, 1: copy, 2: negate, 3: zero
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u) {
        const int64_t t = t0 + (int64_t)u * GT_BLOCK;
        act[u] = 0;
        if (t >= d.total) continue;
        int64_t ij, r, i, j;
        split_index(t, d.rest.n_rest, d.split, d.rest.big, ij, r);
        split_index(ij, d.n, d.split_n, d.rest.big, i, j);
        const bool stored = d.upper ? j >= i : j <= i;
        if (!stored && d.symm == RC_SYMM_N) continue;
        int64_t ro, ri;
        rest_offsets(d.rest, r, ro, ri);
        const int64_t p = stored ? packed_index(d, i, j) : packed_index(d, j, i);
        of[u] = ro + i * d.si + j * d.sj;
        op[u] = ri + p * d.sp;
        act[u] = (anti && i == j) ? 3 : ((anti && !stored) ? 2 : 1);
    }
    Pack<T, V> v[GT_ITEMS];
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u)
        if (act[u] == 1 || act[u] == 2) v[u] = *reinterpret_cast<const Pack<T, V> *>(packed + op[u]);
#pragma unroll
    for (int u = 0; u < GT_ITEMS; ++u) {
        if (act[u] == 0) continue;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (act[u] == 2) v[u].v[k] = -v[u].v[k];
            if (act[u] == 3) v[u].v[k] = T{};
        }
        *reinterpret_cast<Pack<T, V> *>(full + of[u]) = v[u];
    }
}

// ---------------- host side ----------------
void check_dev_ptr(const void *p, const char *name) {
    if (!p) raise(RC_ERR_INVALID_VALUE, std::string("null device pointer: ") + name);
}

Layout without_axes(const Layout &l, int first, int count) {
    Layout r;
    r.offset = l.offset;
    for (int i = 0; i < l.ndim(); ++i)
        if (i < first || i >= first + count) { r.shape.push_back(l.shape[i]); r.stride.push_back(l.stride[i]); }
    return r;
}

// Joint canonical form of the rest dims; returns false if there is nothing to do (zero elements).
bool make_rest(const Layout &lo_rest, const Layout &li_rest, RestDesc *d, int64_t *base_out, int64_t *base_in) {
    RC_CHECK(lo_rest.shape == li_rest.shape, RC_ERR_INVALID_LAYOUT, "Input and output shapes do not match.");
    std::vector<const Layout *> ls{&lo_rest, &li_rest};
    CanonEw c = canon_elementwise(ls, false);
    std::memset(d, 0, sizeof(*d));
    if (c.empty) return false;
    RC_CHECK(c.ndim <= KMAXD, RC_ERR_UNIMPLEMENTED, "more than 8 non-mergeable batch axes");
    d->nd = c.ndim;
    d->n_rest = 1;
    for (int i = 0; i < c.ndim; ++i) {
        d->shape[i] = c.shape[i];
        d->s_out[i] = c.stride[0][i];
        d->s_in[i] = c.stride[1][i];
        if (c.shape[i] >= (1ll << 31)) d->big = 1; else d->dv[i] = FastDiv((uint32_t)c.shape[i]);
        d->n_rest *= c.shape[i];
    }
    if (d->n_rest >= (1ll << 31)) d->big = 1;
    *base_out = c.base[0];
    *base_in = c.base[1];
    return true;
}

// Moves whole `w`-byte words instead of `e`-byte elements when the fastest rest dim is contiguous in both operands
// and every other stride, the extent and both base addresses are multiples of the word.
int promote_word(RestDesc *r, int e, const std::vector<int64_t *> &other_strides, const void *po, const void *pi,
                 int64_t *base_out, int64_t *base_in) {
    if (r->nd == 0 || r->s_out[0] != 1 || r->s_in[0] != 1) return e;
    for (int w = 32; w > e; w /= 2) {
        const int f = w / e;
        bool ok = r->shape[0] % f == 0 && *base_out % f == 0 && *base_in % f == 0 &&
                  reinterpret_cast<uintptr_t>(po) % w == 0 && reinterpret_cast<uintptr_t>(pi) % w == 0;
        for (int i = 1; i < r->nd && ok; ++i) ok = r->s_out[i] % f == 0 && r->s_in[i] % f == 0;
        for (int64_t *s : other_strides) ok = ok && (*s % f == 0);
        if (!ok) continue;
        r->shape[0] /= f;
        r->n_rest /= f;
        if (r->shape[0] < (1ll << 31)) r->dv[0] = FastDiv((uint32_t)std::max<int64_t>(r->shape[0], 1));
        for (int i = 1; i < r->nd; ++i) { r->s_out[i] /= f; r->s_in[i] /= f; }
        for (int64_t *s : other_strides) *s /= f;
        *base_out /= f;
        *base_in /= f;
        return w;
    }
    return e;
}

unsigned grid_for(int64_t total) {
    int64_t g = (total + GT_BLOCK * GT_ITEMS - 1) / (GT_BLOCK * GT_ITEMS);
    RC_CHECK(g < (1ll << 31), RC_ERR_UNIMPLEMENTED, "more than 2^41 elements in one gather launch");
    return (unsigned)g;
}

template <class U>
void launch_select(rc_device *dev, const SelDesc &d, void *c, int64_t bc, const void *a, int64_t ba, const int64_t *idx) {
    index_select_kernel<U><<<grid_for(d.total), GT_BLOCK, 0, dev->stream>>>(d, static_cast<U *>(c) + bc,
                                                                           static_cast<const U *>(a) + ba, idx);
    after_launch(dev, "index_select_kernel");
}

constexpr int64_t SEL_SMEM_MAX = 64 * 1024;  // 3 CTAs per SM

template <class U>
void launch_select_smem(rc_device *dev, const SelDesc &d, void *c, int64_t bc, const void *a, int64_t ba, const int64_t *idx) {
    const size_t smem = (size_t)d.n_src * sizeof(U);
    if (smem > 48 * 1024)
        RC_CUDA(cudaFuncSetAttribute(index_select_smem_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    index_select_smem_kernel<U><<<(unsigned)d.rest.n_rest, GT_BLOCK, smem, dev->stream>>>(d, static_cast<U *>(c) + bc,
                                                                                         static_cast<const U *>(a) + ba, idx);
    after_launch(dev, "index_select_smem_kernel");
}

constexpr int SEL_WIN_W = 2048;                 // outputs per CTA of the windowed gather
constexpr int64_t SEL_WIN_SPAN_BYTES = 48 * 1024;  // largest source span one window may touch

template <class U>
void launch_select_window(rc_device *dev, const SelDesc &d, void *c, int64_t bc, const void *a, int64_t ba, const int64_t *idx,
                          const int64_t *win, int64_t n_chunks, int64_t max_span) {
    const size_t smem = (size_t)max_span * sizeof(U);
    if (smem > 48 * 1024)
        RC_CUDA(cudaFuncSetAttribute(index_select_window_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    index_select_window_kernel<U><<<(unsigned)(d.rest.n_rest * n_chunks), GT_BLOCK, smem, dev->stream>>>(
        d, static_cast<U *>(c) + bc, static_cast<const U *>(a) + ba, idx, win, SEL_WIN_W, FastDiv((uint32_t)n_chunks));
    after_launch(dev, "index_select_window_kernel");
}

template <class U>
void launch_pack(rc_device *dev, const TriMoveDesc &d, void *p, int64_t bp, const void *f, int64_t bf) {
    pack_tri_kernel<U><<<grid_for(d.total), GT_BLOCK, 0, dev->stream>>>(d, static_cast<U *>(p) + bp,
                                                                       static_cast<const U *>(f) + bf);
    after_launch(dev, "pack_tri_kernel");
}

template <class U>
void launch_pack_tile(rc_device *dev, const TriMoveDesc &d, void *p, int64_t bp, const void *f, int64_t bf) {
    const int64_t g = d.ntile_tri * d.rest.n_rest;
    RC_CHECK(g < (1ll << 31), RC_ERR_UNIMPLEMENTED, "pack_tri grid too large");
    if (tri_tile() == 32)
        pack_tri_tile_kernel<U, 32><<<(unsigned)g, GT_BLOCK, 0, dev->stream>>>(d, static_cast<U *>(p) + bp,
                                                                              static_cast<const U *>(f) + bf);
    else
        pack_tri_tile_kernel<U, 64><<<(unsigned)g, GT_BLOCK, 0, dev->stream>>>(d, static_cast<U *>(p) + bp,
                                                                              static_cast<const U *>(f) + bf);
    after_launch(dev, "pack_tri_tile_kernel");
}

template <class T>
void launch_unpack(rc_device *dev, const TriMoveDesc &d, void *f, int64_t bf, const void *p, int64_t bp) {
    const int64_t g = d.ntile_tri * d.rest.n_rest;
    RC_CHECK(g < (1ll << 31), RC_ERR_UNIMPLEMENTED, "unpack_tri grid too large");
    if (unpack_tile() == 32)
        unpack_tri_kernel<T, 32><<<(unsigned)g, GT_BLOCK, 0, dev->stream>>>(d, static_cast<T *>(f) + bf,
                                                                           static_cast<const T *>(p) + bp);
    else
        unpack_tri_kernel<T, 64><<<(unsigned)g, GT_BLOCK, 0, dev->stream>>>(d, static_cast<T *>(f) + bf,
                                                                           static_cast<const T *>(p) + bp);
    after_launch(dev, "unpack_tri_kernel");
}

// Launches the rest-fastest unpack with the widest pack the layouts allow.  `d.rest` dim 0 is contiguous in both operands.
template <class T>
void launch_unpack_rest(rc_device *dev, TriMoveDesc d, void *f, int64_t bf, const void *p, int64_t bp) {
    int V = 32 / (int)sizeof(T);
    for (; V > 1; V /= 2) {
        bool ok = d.rest.shape[0] % V == 0 && bf % V == 0 && bp % V == 0 && d.sp % V == 0 && d.si % V == 0 && d.sj % V == 0 &&
                  reinterpret_cast<uintptr_t>(f) % (V * sizeof(T)) == 0 && reinterpret_cast<uintptr_t>(p) % (V * sizeof(T)) == 0;
        for (int i = 1; i < d.rest.nd && ok; ++i) ok = d.rest.s_out[i] % V == 0 && d.rest.s_in[i] % V == 0;
        if (ok) break;
    }
    d.rest.shape[0] /= V;
    d.rest.n_rest /= V;
    d.rest.s_out[0] = V;
    d.rest.s_in[0] = V;
    if (d.rest.shape[0] < (1ll << 31)) d.rest.dv[0] = FastDiv((uint32_t)std::max<int64_t>(d.rest.shape[0], 1));
    d.total = d.n * d.n * d.rest.n_rest;
    if (d.total >= (1ll << 31)) d.rest.big = 1;
    else { d.split = FastDiv((uint32_t)d.rest.n_rest); d.split_n = FastDiv((uint32_t)d.n); }
    T *fo = static_cast<T *>(f) + bf;
    const T *pi = static_cast<const T *>(p) + bp;
    const unsigned g = grid_for(d.total);
    if (V * sizeof(T) == 32) unpack_tri_rest_kernel<T, 32 / sizeof(T)><<<g, GT_BLOCK, 0, dev->stream>>>(d, fo, pi);
    else if (V * sizeof(T) == 16) unpack_tri_rest_kernel<T, 16 / sizeof(T)><<<g, GT_BLOCK, 0, dev->stream>>>(d, fo, pi);
    else if (V == 2) unpack_tri_rest_kernel<T, 2><<<g, GT_BLOCK, 0, dev->stream>>>(d, fo, pi);
    else unpack_tri_rest_kernel<T, 1><<<g, GT_BLOCK, 0, dev->stream>>>(d, fo, pi);
    after_launch(dev, "unpack_tri_rest_kernel");
}

struct alignas(16) Word16 { uint64_t a, b; };
struct alignas(32) Word32 { uint64_t a, b, c, d; };

// row-major view of the problem: the reference runs col-major devices on reversed axes with the other triangle
// (device_cpu_serial/operators/op_tri.rs:17-26)
void to_row_major(rc_order order, Layout *l1, Layout *l2, rc_uplo *uplo) {
    if (order == RC_COL_MAJOR) {
        *l1 = reversed_axes(*l1);
        *l2 = reversed_axes(*l2);
        *uplo = (*uplo == RC_UPLO_U) ? RC_UPLO_L : RC_UPLO_U;
    }
}

}  // namespace
}  // namespace rc

using namespace rc;

extern "C" {

int rc_index_select(rc_device *dev, rc_dtype t, void *c, const rc_layout *lc_, const void *a, const rc_layout *la_,
                    int axis, const int64_t *indices, int64_t n_indices) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout lc = from_c(lc_), la = from_c(la_);
        const int ndim = lc.ndim();
        RC_CHECK(ndim == la.ndim(), RC_ERR_INVALID_LAYOUT, "Input and output ndim should same.");
        RC_CHECK(axis >= 0 && axis < ndim, RC_ERR_INVALID_VALUE, "axis out of bounds");
        RC_CHECK(n_indices >= 0 && lc.shape[axis] == n_indices, RC_ERR_INVALID_LAYOUT, "Invalid index length.");
        RC_CHECK(n_indices == 0 || indices, RC_ERR_INVALID_VALUE, "null indices");
        int64_t mx = 0;
        for (int64_t i = 0; i < n_indices; ++i) {
            RC_CHECK(indices[i] >= 0, RC_ERR_INDEX, "Index out of range.");
            mx = std::max(mx, indices[i]);
        }
        RC_CHECK(n_indices == 0 || mx < la.shape[axis], RC_ERR_INDEX, "Index out of range.");
        Layout lc_rest = without_axes(lc, axis, 1), la_rest = without_axes(la, axis, 1);
        SelDesc d;
        std::memset(&d, 0, sizeof(d));
        int64_t bc = 0, ba = 0;
        if (n_indices == 0 || !make_rest(lc_rest, la_rest, &d.rest, &bc, &ba)) return;
        check_dev_ptr(c, "c"); check_dev_ptr(a, "a");
        d.n_idx = n_indices;
        d.sc_axis = lc.stride[axis];
        d.sa_axis = la.stride[axis];
        d.n_src = la.shape[axis];
        RC_CHECK(d.sc_axis != 0 || n_indices == 1, RC_ERR_INVALID_LAYOUT, "output layout is broadcast along the indexed axis");
        const int64_t asc = d.sc_axis < 0 ? -d.sc_axis : d.sc_axis;
        d.idx_fastest = (d.rest.nd == 0 || asc < d.rest.s_out[0]) ? 1 : 0;
        int e = (int)dtype_size(t), w = e;
        if (!d.idx_fastest) w = promote_word(&d.rest, e, {&d.sc_axis, &d.sa_axis}, c, a, &bc, &ba);
        d.total = d.rest.n_rest * d.n_idx;
        if (d.total >= (1ll << 31)) d.rest.big = 1;
        else d.split = FastDiv((uint32_t)(d.idx_fastest ? d.n_idx : d.rest.n_rest));

        // staged-row path: gather along the input's contiguous axis, most of each row wanted, rows fit in smem
        const bool staged = d.idx_fastest && d.sa_axis == 1 && d.n_src * e <= SEL_SMEM_MAX && 2 * n_indices >= d.n_src &&
                            d.n_src >= 256 && d.rest.n_rest >= dev->sm_count && d.rest.n_rest < (1ll << 31) &&
                            n_indices < (1ll << 31);
        // windowed path: the same gather on longer rows when every window of SEL_WIN_W indices touches a short source span
        std::vector<int64_t> win;
        int64_t n_chunks = 0, max_span = 0;
        static const bool window_on = [] { const char *v = getenv("RC_SEL_WINDOW"); return !(v && v[0] == '0'); }();
        if (window_on && !staged && d.idx_fastest && d.sa_axis == 1 && n_indices >= 4 * SEL_WIN_W && n_indices < (1ll << 31) && e <= 8) {
            n_chunks = (n_indices + SEL_WIN_W - 1) / SEL_WIN_W;
            win.resize((size_t)n_chunks * 2);
            int64_t total_span = 0;
            bool ok = d.rest.n_rest * n_chunks < (1ll << 31);
            for (int64_t k = 0; k < n_chunks && ok; ++k) {
                const int64_t i0 = k * SEL_WIN_W, i1 = std::min<int64_t>(n_indices, i0 + SEL_WIN_W);
                int64_t lo = indices[i0], hi = indices[i0];
                for (int64_t i = i0 + 1; i < i1; ++i) { lo = std::min(lo, indices[i]); hi = std::max(hi, indices[i]); }
                const int64_t span = hi - lo + 1;
                ok = span * e <= SEL_WIN_SPAN_BYTES;
                win[2 * k] = lo;
                win[2 * k + 1] = span;
                total_span += span;
                max_span = std::max(max_span, span);
            }
            // Dense clustered selections only (at least ~5 of 8 source elements wanted): measured on (2048, 32768) f64, a 70 %
            // mask gains (3.2 -> 3.8 TB/s) while "every other column" loses (3.2 -> 2.9): with half of every sector wanted
            // the direct kernel's sector reads already share sectors between neighbours and move no more than the window.
            if (!ok || 5 * total_span > 8 * n_indices) n_chunks = 0;
        }

        int64_t *idx_dev = nullptr;
        cudaError_t err = cudaMallocAsync(reinterpret_cast<void **>(&idx_dev), (size_t)(n_indices + 2 * n_chunks) * 8, dev->stream);
        if (err != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(err));
        try {
            upload_small(dev, idx_dev, indices, (size_t)n_indices * 8);  // pinned ring: no stream sync between calls
            if (n_chunks > 0) {
                upload_small(dev, idx_dev + n_indices, win.data(), (size_t)n_chunks * 16);
                w = -100 - e;
            }
            if (staged) w = -e;
            switch (w) {
                case -101: launch_select_window<uint8_t>(dev, d, c, bc, a, ba, idx_dev, idx_dev + n_indices, n_chunks, max_span); break;
                case -102: launch_select_window<uint16_t>(dev, d, c, bc, a, ba, idx_dev, idx_dev + n_indices, n_chunks, max_span); break;
                case -104: launch_select_window<uint32_t>(dev, d, c, bc, a, ba, idx_dev, idx_dev + n_indices, n_chunks, max_span); break;
                case -108: launch_select_window<uint64_t>(dev, d, c, bc, a, ba, idx_dev, idx_dev + n_indices, n_chunks, max_span); break;
                case -1: launch_select_smem<uint8_t>(dev, d, c, bc, a, ba, idx_dev); break;
                case -2: launch_select_smem<uint16_t>(dev, d, c, bc, a, ba, idx_dev); break;
                case -4: launch_select_smem<uint32_t>(dev, d, c, bc, a, ba, idx_dev); break;
                case -8: launch_select_smem<uint64_t>(dev, d, c, bc, a, ba, idx_dev); break;
                case 1: launch_select<uint8_t>(dev, d, c, bc, a, ba, idx_dev); break;
                case 2: launch_select<uint16_t>(dev, d, c, bc, a, ba, idx_dev); break;
                case 4: launch_select<uint32_t>(dev, d, c, bc, a, ba, idx_dev); break;
                case 8: launch_select<uint64_t>(dev, d, c, bc, a, ba, idx_dev); break;
                case 16: launch_select<Word16>(dev, d, c, bc, a, ba, idx_dev); break;
                default: launch_select<Word32>(dev, d, c, bc, a, ba, idx_dev); break;
            }
        } catch (...) {
            cudaFreeAsync(idx_dev, dev->stream);
            throw;
        }
        RC_CUDA(cudaFreeAsync(idx_dev, dev->stream));
    });
}

int rc_pack_tri(rc_device *dev, rc_dtype t, void *a, const rc_layout *la_, const void *b, const rc_layout *lb_, rc_uplo uplo) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout la = from_c(la_), lb = from_c(lb_);
        RC_CHECK(uplo == RC_UPLO_U || uplo == RC_UPLO_L, RC_ERR_INVALID_VALUE, "uplo must be RC_UPLO_U or RC_UPLO_L");
        RC_CHECK(lb.ndim() >= 2 && la.ndim() + 1 == lb.ndim(), RC_ERR_INVALID_LAYOUT,
                 "pack_tri: the packed tensor has one axis less than the full one");
        to_row_major(dev->order, &la, &lb, &uplo);
        const int nb = lb.ndim();
        const int64_t n = lb.shape[nb - 1];
        RC_CHECK(lb.shape[nb - 2] == n, RC_ERR_INVALID_LAYOUT, "Last two dimensions should be the same for pack_tri.");
        RC_CHECK(la.shape[nb - 2] == n * (n + 1) / 2, RC_ERR_INVALID_LAYOUT, "pack_tri: packed axis must have n (n + 1) / 2 elements");
        TriMoveDesc d;
        std::memset(&d, 0, sizeof(d));
        int64_t bp = 0, bf = 0;
        Layout la_rest = without_axes(la, nb - 2, 1), lb_rest = without_axes(lb, nb - 2, 2);
        if (n == 0 || !make_rest(la_rest, lb_rest, &d.rest, &bp, &bf)) return;
        check_dev_ptr(a, "a"); check_dev_ptr(b, "b");
        d.n = n;
        d.n_tp = n * (n + 1) / 2;
        d.sp = la.stride[nb - 2];
        d.si = lb.stride[nb - 2];
        d.sj = lb.stride[nb - 1];
        d.upper = uplo == RC_UPLO_U;
        RC_CHECK(d.sp != 0 || d.n_tp == 1, RC_ERR_INVALID_LAYOUT, "output layout is broadcast along the packed axis");
        const int64_t asp = d.sp < 0 ? -d.sp : d.sp;
        d.tri_fastest = (d.rest.nd == 0 || asp < d.rest.s_out[0]) ? 1 : 0;
        int e = (int)dtype_size(t), w = e;
        if (d.tri_fastest) {  // tile kernel: packed rows and full rows both move in row segments
            d.ntile = (n + tri_tile() - 1) / tri_tile();
            d.ntile_tri = d.ntile * (d.ntile + 1) / 2;
            switch (e) {
                case 1: launch_pack_tile<uint8_t>(dev, d, a, bp, b, bf); break;
                case 2: launch_pack_tile<uint16_t>(dev, d, a, bp, b, bf); break;
                case 4: launch_pack_tile<uint32_t>(dev, d, a, bp, b, bf); break;
                default: launch_pack_tile<uint64_t>(dev, d, a, bp, b, bf); break;
            }
            return;
        }
        w = promote_word(&d.rest, e, {&d.sp, &d.si, &d.sj}, a, b, &bp, &bf);
        d.total = d.rest.n_rest * d.n_tp;
        if (d.total >= (1ll << 31)) d.rest.big = 1;
        else d.split = FastDiv((uint32_t)(d.tri_fastest ? d.n_tp : d.rest.n_rest));
        switch (w) {
            case 1: launch_pack<uint8_t>(dev, d, a, bp, b, bf); break;
            case 2: launch_pack<uint16_t>(dev, d, a, bp, b, bf); break;
            case 4: launch_pack<uint32_t>(dev, d, a, bp, b, bf); break;
            case 8: launch_pack<uint64_t>(dev, d, a, bp, b, bf); break;
            case 16: launch_pack<Word16>(dev, d, a, bp, b, bf); break;
            default: launch_pack<Word32>(dev, d, a, bp, b, bf); break;
        }
    });
}

int rc_unpack_tri(rc_device *dev, rc_dtype t, void *a, const rc_layout *la_, const void *b, const rc_layout *lb_,
                  rc_uplo uplo, rc_symm symm) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout la = from_c(la_), lb = from_c(lb_);
        RC_CHECK(uplo == RC_UPLO_U || uplo == RC_UPLO_L, RC_ERR_INVALID_VALUE, "uplo must be RC_UPLO_U or RC_UPLO_L");
        RC_CHECK((int)symm >= RC_SYMM_SY && (int)symm <= RC_SYMM_N, RC_ERR_INVALID_VALUE, "unknown symm flag");
        RC_CHECK(dtype_is_float(t), RC_ERR_UNIMPLEMENTED, "unpack_tri is defined for floating-point element types");
        RC_CHECK(la.ndim() >= 2 && lb.ndim() + 1 == la.ndim(), RC_ERR_INVALID_LAYOUT,
                 "unpack_tri: the packed tensor has one axis less than the full one");
        to_row_major(dev->order, &la, &lb, &uplo);
        const int na = la.ndim();
        const int64_t n = la.shape[na - 1];
        RC_CHECK(la.shape[na - 2] == n, RC_ERR_INVALID_LAYOUT, "Last two dimensions should be the same for unpack_tri.");
        RC_CHECK(lb.shape[na - 2] == n * (n + 1) / 2, RC_ERR_INVALID_LAYOUT,
                 "Last dimension should be triangular number for unpack_tri.");
        TriMoveDesc d;
        std::memset(&d, 0, sizeof(d));
        int64_t bf = 0, bp = 0;
        Layout la_rest = without_axes(la, na - 2, 2), lb_rest = without_axes(lb, na - 2, 1);
        if (n == 0 || !make_rest(la_rest, lb_rest, &d.rest, &bf, &bp)) return;
        check_dev_ptr(a, "a"); check_dev_ptr(b, "b");
        d.n = n;
        d.n_tp = n * (n + 1) / 2;
        d.sp = lb.stride[na - 2];
        d.si = la.stride[na - 2];
        d.sj = la.stride[na - 1];
        d.upper = uplo == RC_UPLO_U;
        d.symm = (int)symm;
        RC_CHECK((d.si != 0 && d.sj != 0) || n == 1, RC_ERR_INVALID_LAYOUT, "output layout is broadcast along the matrix axes");
        const int64_t asi = d.si < 0 ? -d.si : d.si, asj = d.sj < 0 ? -d.sj : d.sj;
        if (d.rest.nd > 0 && d.rest.s_out[0] == 1 && d.rest.s_in[0] == 1 && asi != 1 && asj != 1 && d.rest.shape[0] >= 4) {
            // the batch axis is the contiguous one: move runs of it
            if (t == RC_F64) launch_unpack_rest<double>(dev, d, a, bf, b, bp);
            else launch_unpack_rest<float>(dev, d, a, bf, b, bp);
            return;
        }
        d.ntile = (n + unpack_tile() - 1) / unpack_tile();
        d.ntile_tri = d.ntile * (d.ntile + 1) / 2;
        if (t == RC_F64) launch_unpack<double>(dev, d, a, bf, b, bp);
        else launch_unpack<float>(dev, d, a, bf, b, bp);
    });
}

}  // extern "C"
