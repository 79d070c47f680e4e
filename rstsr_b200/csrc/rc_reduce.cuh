// rc_reduce.cuh -- kernel families K3/K4: sum / prod / max / min / mean over all or selected axes.
//
// Replaces rstsr-native-impl/src/cpu_rayon/reduction.rs:20-328 (reduce_all_cpu_rayon, reduce_axes_cpu_rayon)
// with the (init, f, f_sum, f_out) monoids of rstsr-core/src/feature_rayon/auto_impl/reduction.rs:7-205:
//   sum: 0, +      prod: 1, *      max: T::MIN, ext_max      min: T::MAX, ext_min      mean: sum, then / n
// Float max/min ignore NaN and start from the finite extreme (rstsr-dtype-traits/src/ext_real.rs:70-87);
// the accumulator has the element type (f32 sums in f32); integer sums wrap.
//
//   reduce_rows_kernel  a group of G threads (1..256) owns one output and strides over the reduced index
//                       space; when the smallest-stride reduced dim is contiguous it is read as 32-byte packs
//                       (256-bit LDG), 8 packs in flight per thread.  Warp-shuffle tree inside a warp,
//                       shared-memory tree across warps.
//   reduce_cols_kernel  the kept fastest axis is contiguous in input and output: lanes own columns
//                       (32-byte packs), warps walk the reduced rows: coalesced 1 KiB per warp and row.
// Both take a split factor S along the reduced space (gridDim.y); S > 1 writes partials[S][n_out] into the
// device handle's workspace and a second launch of the same kernel folds them in a FIXED order: the result
// is run-to-run deterministic (no float atomics), which the reference's rayon fold is not.
//
// Measured on B200 (scripts/membench.cu): a read-only stream needs >= 128 KB in flight per SM to reach
// 6.7-7.0 TB/s; 256-bit loads x 8 per thread do, 128-bit x 4 stall at 6.1 TB/s.
#pragma once
#include <limits>

#include "rc_kernel_common.cuh"
#include "rc_ops.hpp"

namespace rc {

namespace {

constexpr int RED_BLOCK = 256;
constexpr int RED_UNROLL = 8;

struct RedDesc {
    int nk, nr;
    int big;                    // 1: some index does not fit the 32-bit fast-division path
    int64_t kshape[KMAXD], ks_in[KMAXD], ks_out[KMAXD];
    int64_t rshape[KMAXD], rs[KMAXD];
    FastDiv kdiv[KMAXD], rdiv[KMAXD];
    int64_t n_out;              // outputs (rows kernel) / outputs excluding kept dim 0 (cols kernel)
    int64_t n_items;            // reduced index space (dim 0 in packs for the vectorised rows kernel)
    int64_t chunk;              // reduced items per split
    int64_t n_out_total;        // elements of the output (partial row pitch)
    int64_t packs0;             // cols kernel: packs along kept dim 0
    int group;                  // rows kernel: threads per output
    int tcol;                   // cols kernel: threads along the kept axis
    int do_div;                 // mean: divide by div on the final write
    int to_partial;             // write un-finalised accumulators to the partial buffer
};

template <class T> struct OpSum {
    static __device__ __forceinline__ T init() { return (T)0; }
    static __device__ __forceinline__ T f(T a, T b) {
        if constexpr (std::is_integral<T>::value) return (T)((typename std::make_unsigned<T>::type)a + (typename std::make_unsigned<T>::type)b);
        else return a + b;
    }
};
template <class T> struct OpProd {
    static __device__ __forceinline__ T init() { return (T)1; }
    static __device__ __forceinline__ T f(T a, T b) {
        if constexpr (std::is_integral<T>::value) return (T)((typename std::make_unsigned<T>::type)a * (typename std::make_unsigned<T>::type)b);
        else return a * b;
    }
};
template <class T> struct OpMax {
    static __device__ __forceinline__ T init() { return std::numeric_limits<T>::lowest(); }
    static __device__ __forceinline__ T f(T a, T b) {
        if constexpr (std::is_same<T, float>::value) return fmaxf(a, b);
        else if constexpr (std::is_same<T, double>::value) return fmax(a, b);
        else return a < b ? b : a;
    }
};
template <class T> struct OpMin {
    static __device__ __forceinline__ T init() { return std::numeric_limits<T>::max(); }
    static __device__ __forceinline__ T f(T a, T b) {
        if constexpr (std::is_same<T, float>::value) return fminf(a, b);
        else if constexpr (std::is_same<T, double>::value) return fmin(a, b);
        else return b < a ? b : a;
    }
};

template <class T>
__device__ __forceinline__ T finalize(T acc, int do_div, T div) {
    if constexpr (std::is_floating_point<T>::value) {
        if (do_div) return acc / div;
    }
    return acc;
}

// offset of linear index `i` over dims [first, n) with the given strides (cold path: kept never inlined in
// the streaming loop unless the reduced space really is multi-dimensional)
__device__ __noinline__ int64_t decompose_big(int64_t i, int first, int n, const int64_t *shape, const int64_t *stride) {
    int64_t off = 0;
    for (int k = first; k < n; ++k) {
        int64_t q, r;
        if (k + 1 < n) { q = i / shape[k]; r = i - q * shape[k]; } else { q = 0; r = i; }
        off += r * stride[k];
        i = q;
    }
    return off;
}

__device__ __forceinline__ int64_t decompose(int64_t i, int first, int n, const int64_t *shape, const FastDiv *dv,
                                             const int64_t *stride, int big) {
    if (big) return decompose_big(i, first, n, shape, stride);
    int64_t off = 0;
    uint32_t t = (uint32_t)i;
#pragma unroll 1
    for (int k = first; k < n; ++k) {
        uint32_t q, r;
        if (k + 1 < n) dv[k].divmod(t, q, r); else { q = 0; r = t; }
        off += (int64_t)r * stride[k];
        t = q;
    }
    return off;
}

template <class T>
__device__ __forceinline__ T shfl_xor_t(T v, int m) {
    if constexpr (sizeof(T) == 8) {
        long long x = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<long long *>(&v), m);
        return *reinterpret_cast<T *>(&x);
    } else {
        int x = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<int *>(&v), m);
        return *reinterpret_cast<T *>(&x);
    }
}

// ------------------------------------------------------------------------------------------------
// MULTI: the reduced index space has more than one (non-mergeable) dim -> per-item decomposition.
template <class Op, class T, int VEC, bool MULTI>
__global__ void __launch_bounds__(RED_BLOCK) reduce_rows_kernel(const __grid_constant__ RedDesc d, const T *__restrict__ in,
                                                                T *__restrict__ out, T *__restrict__ partial,
                                                                T div) {
    __shared__ T warp_acc[RED_BLOCK / 32];
    const int G = d.group;
    const int tid = threadIdx.x;
    const int g = tid / G, t = tid - g * G;
    const int64_t o = (int64_t)blockIdx.x * (RED_BLOCK / G) + g;
    const bool valid = o < d.n_out;
    int64_t off_out = 0;
    const T *src = in;
    if (valid) {
        src += decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_in, d.big);
        off_out = decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_out, d.big);
    }
    const int64_t begin = (int64_t)blockIdx.y * d.chunk;
    int64_t end = begin + d.chunk;
    if (end > d.n_items) end = d.n_items;
    if (!valid) end = begin;

    T acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = Op::init();

    const int64_t rs0 = d.rs[0];  // per item (already multiplied by VEC for packs)
    int64_t i = begin + t;
    // full groups: RED_UNROLL independent loads in flight, no per-load predicate
    for (; i + (int64_t)(RED_UNROLL - 1) * G < end; i += (int64_t)G * RED_UNROLL) {
        Pack<T, VEC> p[RED_UNROLL];
#pragma unroll
        for (int u = 0; u < RED_UNROLL; ++u) {
            const int64_t ii = i + (int64_t)u * G;
            int64_t off;
            if constexpr (MULTI) off = decompose(ii, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
            else off = ii * rs0;
            p[u] = ld_stream<T, VEC>(src + off);
        }
#pragma unroll
        for (int u = 0; u < RED_UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = Op::f(acc[j], p[u].v[j]);
    }
    for (; i < end; i += G) {
        int64_t off;
        if constexpr (MULTI) off = decompose(i, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
        else off = i * rs0;
        Pack<T, VEC> p = ld_stream<T, VEC>(src + off);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = Op::f(acc[j], p.v[j]);
    }
    // fold the pack lanes, then the group, in a fixed order
    T v = acc[0];
#pragma unroll
    for (int j = 1; j < VEC; ++j) v = Op::f(v, acc[j]);

    if (G <= 32) {
        for (int m = G >> 1; m >= 1; m >>= 1) v = Op::f(v, shfl_xor_t<T>(v, m));
    } else {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v = Op::f(v, shfl_xor_t<T>(v, m));
        if ((tid & 31) == 0) warp_acc[tid >> 5] = v;
        __syncthreads();
        if (t == 0) {
            const int w0 = tid >> 5, nw = G >> 5;
            v = warp_acc[w0];
            for (int w = 1; w < nw; ++w) v = Op::f(v, warp_acc[w0 + w]);
        }
    }
    if (valid && t == 0) {
        if (d.to_partial) partial[(int64_t)blockIdx.y * d.n_out_total + o] = v;
        else out[off_out] = finalize<T>(v, d.do_div, div);
    }
}

// ------------------------------------------------------------------------------------------------
template <class Op, class T, int VEC, bool MULTI>
__global__ void __launch_bounds__(RED_BLOCK) reduce_cols_kernel(const __grid_constant__ RedDesc d, const T *__restrict__ in,
                                                                T *__restrict__ out, T *__restrict__ partial,
                                                                T div) {
    extern __shared__ __align__(32) unsigned char red_smem[];
    Pack<T, VEC> *sm = reinterpret_cast<Pack<T, VEC> *>(red_smem);
    const int TC = d.tcol, RW = RED_BLOCK / TC;
    const int tx = threadIdx.x % TC, ty = threadIdx.x / TC;
    const int64_t ntile0 = (d.packs0 + TC - 1) / TC;
    const int64_t tile = blockIdx.x;
    const int64_t kb = tile / ntile0;                    // linear index over kept dims 1..
    const int64_t col = (tile - kb * ntile0) * TC + tx;  // pack index along kept dim 0
    const bool valid = col < d.packs0;
    int64_t off_out = 0;
    const T *src = in;
    if (valid) {
        src += col * VEC + decompose(kb, 1, d.nk, d.kshape, d.kdiv, d.ks_in, d.big);
        off_out = col * VEC + decompose(kb, 1, d.nk, d.kshape, d.kdiv, d.ks_out, d.big);
    }
    const int64_t begin = (int64_t)blockIdx.y * d.chunk;
    int64_t end = begin + d.chunk;
    if (end > d.n_items) end = d.n_items;
    if (!valid) end = begin;

    T acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = Op::init();

    const int64_t rs0 = d.rs[0];
    int64_t r = begin + ty;
    for (; r + (int64_t)(RED_UNROLL - 1) * RW < end; r += (int64_t)RW * RED_UNROLL) {
        Pack<T, VEC> p[RED_UNROLL];
#pragma unroll
        for (int u = 0; u < RED_UNROLL; ++u) {
            const int64_t rr = r + (int64_t)u * RW;
            int64_t off;
            if constexpr (MULTI) off = decompose(rr, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
            else off = rr * rs0;
            p[u] = ld_stream<T, VEC>(src + off);
        }
#pragma unroll
        for (int u = 0; u < RED_UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = Op::f(acc[j], p[u].v[j]);
    }
    for (; r < end; r += RW) {
        int64_t off;
        if constexpr (MULTI) off = decompose(r, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
        else off = r * rs0;
        Pack<T, VEC> p = ld_stream<T, VEC>(src + off);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = Op::f(acc[j], p.v[j]);
    }
    Pack<T, VEC> v;
#pragma unroll
    for (int j = 0; j < VEC; ++j) v.v[j] = acc[j];
    sm[ty * TC + tx] = v;
    __syncthreads();
    if (ty == 0 && valid) {
        for (int w = 1; w < RW; ++w) {
            Pack<T, VEC> o = sm[w * TC + tx];
#pragma unroll
            for (int j = 0; j < VEC; ++j) v.v[j] = Op::f(v.v[j], o.v[j]);
        }
        if (d.to_partial) {
            T *dst = partial + (int64_t)blockIdx.y * d.n_out_total + kb * (d.packs0 * VEC) + col * VEC;
            if (VEC > 1) st_stream<T, VEC>(dst, v);
            else dst[0] = v.v[0];
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) v.v[j] = finalize<T>(v.v[j], d.do_div, div);
            if (VEC > 1) st_stream<T, VEC>(out + off_out, v);
            else out[off_out] = v.v[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void fill_desc_dims(RedDesc &d, const CanonRed &c) {
    RC_CHECK(c.kshape.size() <= (size_t)KMAXD && c.rshape.size() <= (size_t)KMAXD, RC_ERR_UNIMPLEMENTED,
             "reduction over more than 8 non-mergeable kept or reduced axes");
    d.nk = (int)c.kshape.size();
    d.nr = (int)c.rshape.size();
    d.big = 0;
    for (int i = 0; i < d.nk; ++i) {
        d.kshape[i] = c.kshape[i];
        d.ks_in[i] = c.kstride_in[i];
        d.ks_out[i] = c.kstride_out[i];
        if (c.kshape[i] >= (1ll << 31)) d.big = 1; else d.kdiv[i] = FastDiv((uint32_t)c.kshape[i]);
    }
    for (int i = 0; i < d.nr; ++i) {
        d.rshape[i] = c.rshape[i];
        d.rs[i] = c.rstride[i];
        if (c.rshape[i] >= (1ll << 31)) d.big = 1; else d.rdiv[i] = FastDiv((uint32_t)std::max<int64_t>(c.rshape[i], 1));
    }
}

template <class T>
bool aligned_for(const void *p, int vec) { return reinterpret_cast<uintptr_t>(p) % (vec * sizeof(T)) == 0; }

template <class Op, class T>
void launch_rows(rc_device *dev, const RedDesc &d, int vec, int64_t sy, const T *in, T *out, T *partial, T div) {
    int64_t gx = (d.n_out + (RED_BLOCK / d.group) - 1) / (RED_BLOCK / d.group);
    RC_CHECK(gx < (1ll << 31) && sy <= 65535, RC_ERR_UNIMPLEMENTED, "reduction grid too large");
    dim3 grid((unsigned)gx, (unsigned)sy);
    constexpr int V = 32 / sizeof(T);
    const bool multi = d.nr > 1;
    if (vec > 1) {
        if (multi) reduce_rows_kernel<Op, T, V, true><<<grid, RED_BLOCK, 0, dev->stream>>>(d, in, out, partial, div);
        else reduce_rows_kernel<Op, T, V, false><<<grid, RED_BLOCK, 0, dev->stream>>>(d, in, out, partial, div);
    } else {
        if (multi) reduce_rows_kernel<Op, T, 1, true><<<grid, RED_BLOCK, 0, dev->stream>>>(d, in, out, partial, div);
        else reduce_rows_kernel<Op, T, 1, false><<<grid, RED_BLOCK, 0, dev->stream>>>(d, in, out, partial, div);
    }
    after_launch(dev, "reduce_rows_kernel");
}

template <class Op, class T>
void launch_cols(rc_device *dev, const RedDesc &d, int vec, int64_t sy, const T *in, T *out, T *partial, T div) {
    int64_t ntile0 = (d.packs0 + d.tcol - 1) / d.tcol;
    int64_t gx = ntile0 * d.n_out;
    RC_CHECK(gx < (1ll << 31) && sy <= 65535, RC_ERR_UNIMPLEMENTED, "reduction grid too large");
    dim3 grid((unsigned)gx, (unsigned)sy);
    constexpr int V = 32 / sizeof(T);
    size_t smem = (size_t)RED_BLOCK * sizeof(T) * (vec > 1 ? V : 1);
    const bool multi = d.nr > 1;
    if (vec > 1) {
        if (multi) reduce_cols_kernel<Op, T, V, true><<<grid, RED_BLOCK, smem, dev->stream>>>(d, in, out, partial, div);
        else reduce_cols_kernel<Op, T, V, false><<<grid, RED_BLOCK, smem, dev->stream>>>(d, in, out, partial, div);
    } else {
        if (multi) reduce_cols_kernel<Op, T, 1, true><<<grid, RED_BLOCK, smem, dev->stream>>>(d, in, out, partial, div);
        else reduce_cols_kernel<Op, T, 1, false><<<grid, RED_BLOCK, smem, dev->stream>>>(d, in, out, partial, div);
    }
    after_launch(dev, "reduce_cols_kernel");
}

// threads along the kept (contiguous) axis of the column kernel; RC_TCOL_MAX is a tuning knob for experiments
inline int tcol_max() {
    static int v = [] { const char *e = getenv("RC_TCOL_MAX"); int x = e ? atoi(e) : 64; return (x >= 1 && x <= RED_BLOCK) ? x : 64; }();
    return v;
}

int pow2_floor(int64_t x) { int p = 1; while ((int64_t)p * 2 <= x) p *= 2; return p; }
int pow2_ceil(int64_t x) { int p = 1; while (p < x) p *= 2; return p; }

template <class Op, class T>
void reduce_typed(rc_device *dev, const CanonRed &c, const void *a_v, void *out_v, bool mean, int64_t mean_count) {
    if (c.empty_out) return;
    const T *in = static_cast<const T *>(a_v) + c.base_in;
    T *out = static_cast<T *>(out_v) + c.base_out;
    T div = (T)1;
    if (mean) div = (T)mean_count;  // T::from_usize(n) (auto_impl/reduction.rs:181,199)
    constexpr int V = 32 / sizeof(T);
    const int64_t n_out = c.n_out(), n_red = c.n_red();
    // 2 resident CTAs per SM keep 2 x 256 threads x 8 x 32 B = 128 KB in flight; aim for >= 4 waves of work
    // two FULL waves at the kernels' occupancy (4 CTAs of 256 threads per SM at <= 64 registers): a split that
    // overshoots a wave boundary leaves a tail wave (measured: 1280 vs 1184 CTAs costs 4 %)
    const int64_t target_ctas = (int64_t)dev->sm_count * 8;

    RedDesc d;
    std::memset(&d, 0, sizeof(d));
    fill_desc_dims(d, c);
    d.do_div = mean ? 1 : 0;
    d.n_out_total = n_out;
    if (n_out >= (1ll << 31)) d.big = 1;

    const bool red_contig = d.nr >= 1 && d.rs[0] == 1 && d.rshape[0] >= 8;
    const bool kept_contig = d.nk >= 1 && d.ks_in[0] == 1 && d.ks_out[0] == 1 && d.kshape[0] >= 8;

    if (!red_contig && kept_contig && n_red > 0) {
        // ---------------- column kernel ----------------
        bool vec_ok = d.kshape[0] % V == 0 && aligned_for<T>(in, V) && aligned_for<T>(out, V);
        for (int i = 1; i < d.nk && vec_ok; ++i) vec_ok = d.ks_in[i] % V == 0 && d.ks_out[i] % V == 0;
        for (int i = 0; i < d.nr && vec_ok; ++i) vec_ok = d.rs[i] % V == 0;
        const int vec = vec_ok ? V : 1;
        d.packs0 = d.kshape[0] / vec;
        d.tcol = (int)std::min<int64_t>(tcol_max(), pow2_ceil(d.packs0));
        d.n_out = n_out / d.kshape[0];
        d.n_items = n_red;
        if (n_red >= (1ll << 31) || d.n_out >= (1ll << 31)) d.big = 1;
        const int rw = RED_BLOCK / d.tcol;
        int64_t base_ctas = ((d.packs0 + d.tcol - 1) / d.tcol) * d.n_out;
        int64_t S = std::min<int64_t>(std::max<int64_t>(1, target_ctas / base_ctas),
                                      std::max<int64_t>(1, n_red / ((int64_t)rw * RED_UNROLL * 2)));
        S = std::max<int64_t>(1, std::min<int64_t>(S, 1024));
        d.chunk = (n_red + S - 1) / S;
        S = (n_red + d.chunk - 1) / d.chunk;
        if (S == 1) {
            d.to_partial = 0;
            launch_cols<Op, T>(dev, d, vec, 1, in, out, (T *)nullptr, div);
            return;
        }
        T *partial = static_cast<T *>(workspace(dev, (size_t)S * n_out * sizeof(T)));
        d.to_partial = 1;
        launch_cols<Op, T>(dev, d, vec, S, in, out, partial, div);
        // second pass: fold partial[S][n_out] over S, same kept dims with contiguous input strides
        RedDesc e = d;
        e.to_partial = 0;
        e.nr = 1;
        e.rshape[0] = S;
        e.rs[0] = n_out;
        e.rdiv[0] = FastDiv((uint32_t)S);
        int64_t acc = 1;
        for (int i = 0; i < e.nk; ++i) { e.ks_in[i] = acc; acc *= e.kshape[i]; }
        e.n_items = S;
        e.chunk = S;
        const bool vec2 = vec_ok && (n_out % V == 0) && aligned_for<T>(partial, V);
        const int v2 = vec2 ? V : 1;
        e.packs0 = e.kshape[0] / v2;
        e.tcol = (int)std::min<int64_t>(tcol_max(), pow2_ceil(e.packs0));
        launch_cols<Op, T>(dev, e, v2, 1, partial, out, (T *)nullptr, div);
        return;
    }

    // ---------------- row kernel (contiguous or generic reduced space) ----------------
    bool vec_ok = d.nr >= 1 && d.rs[0] == 1 && d.rshape[0] % V == 0 && aligned_for<T>(in, V);
    for (int i = 0; i < d.nk && vec_ok; ++i) vec_ok = d.ks_in[i] % V == 0;
    for (int i = 1; i < d.nr && vec_ok; ++i) vec_ok = d.rs[i] % V == 0;
    const int vec = vec_ok ? V : 1;
    if (d.nr == 0) {  // nothing reduced (all reduced axes have extent 1): a strided copy through the monoid
        d.nr = 1;
        d.rshape[0] = 1;
        d.rs[0] = 0;
        d.rdiv[0] = FastDiv(1);
    }
    if (vec > 1) {
        d.rshape[0] /= vec;
        d.rdiv[0] = FastDiv((uint32_t)std::max<int64_t>(d.rshape[0], 1));
        d.rs[0] = vec;
    }
    d.n_out = n_out;
    d.n_items = n_red / vec;
    if (d.n_items >= (1ll << 31) && d.nr > 1) d.big = 1;
    d.group = (int)std::min<int64_t>(RED_BLOCK, std::max<int64_t>(1, pow2_floor(std::max<int64_t>(1, d.n_items / RED_UNROLL))));
    int64_t base_ctas = (n_out + (RED_BLOCK / d.group) - 1) / (RED_BLOCK / d.group);
    int64_t S = 1;
    if (base_ctas < target_ctas) {
        S = std::min<int64_t>(std::max<int64_t>(1, target_ctas / base_ctas),
                              std::max<int64_t>(1, d.n_items / ((int64_t)d.group * RED_UNROLL * 4)));
        S = std::max<int64_t>(1, std::min<int64_t>(S, 4096));
    }
    d.chunk = (d.n_items + S - 1) / std::max<int64_t>(S, 1);
    if (d.chunk == 0) d.chunk = 1;
    S = std::max<int64_t>(1, (d.n_items + d.chunk - 1) / d.chunk);
    if (S == 1) {
        d.to_partial = 0;
        launch_rows<Op, T>(dev, d, vec, 1, in, out, (T *)nullptr, div);
        return;
    }
    T *partial = static_cast<T *>(workspace(dev, (size_t)S * n_out * sizeof(T)));
    d.to_partial = 1;
    launch_rows<Op, T>(dev, d, vec, S, in, out, partial, div);
    // second pass: out[o] = fold_s partial[s][o]
    RedDesc e = d;
    e.to_partial = 0;
    e.nr = 1;
    e.rshape[0] = S;
    e.rs[0] = n_out;
    e.rdiv[0] = FastDiv((uint32_t)S);
    int64_t acc = 1;
    for (int i = 0; i < e.nk; ++i) { e.ks_in[i] = acc; acc *= e.kshape[i]; }
    e.n_items = S;
    e.chunk = S;
    e.group = (int)std::min<int64_t>(RED_BLOCK, std::max<int64_t>(1, pow2_floor(std::max<int64_t>(1, S / 2))));
    launch_rows<Op, T>(dev, e, 1, 1, partial, out, (T *)nullptr, div);
}

template <class T>
void reduce_op(rc_device *dev, rc_redop op, const CanonRed &c, const void *a, void *out, int64_t mean_count) {
    switch (op) {
        case RC_SUM: reduce_typed<OpSum<T>, T>(dev, c, a, out, false, 1); return;
        case RC_PROD: reduce_typed<OpProd<T>, T>(dev, c, a, out, false, 1); return;
        case RC_MAX: reduce_typed<OpMax<T>, T>(dev, c, a, out, false, 1); return;
        case RC_MIN: reduce_typed<OpMin<T>, T>(dev, c, a, out, false, 1); return;
        case RC_MEAN:
            if constexpr (std::is_floating_point<T>::value) {
                reduce_typed<OpSum<T>, T>(dev, c, a, out, true, mean_count);
                return;
            } else {
                raise(RC_ERR_UNIMPLEMENTED, "mean is only defined for floating-point element types");
            }
    }
    raise(RC_ERR_INVALID_VALUE, "unknown reduction op");
}

}  // namespace
}  // namespace rc
