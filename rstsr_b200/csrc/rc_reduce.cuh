// rc_reduce.cuh -- kernel families K3/K4: reductions over all or selected axes.
//
// Replaces rstsr-native-impl/src/cpu_rayon/reduction.rs:20-328 (reduce_all_cpu_rayon, reduce_axes_cpu_rayon) and
// the arg variants (cpu_serial/reduction.rs:421-582) with the (init, f, f_sum, f_out) monoids of
// rstsr-core/src/feature_rayon/auto_impl/reduction.rs:
//   sum: 0, +      prod: 1, *      max: T::MIN, ext_max      min: T::MAX, ext_min      mean: sum, then / n   (:7-205)
//   var/std: (sum x, sum x^2) -> q/n - (s/n)^2 [, sqrt]      l2_norm: sum x^2 -> sqrt                        (:207-354)
//   argmin/argmax: (value, row-major index), first occurrence wins                                          (:356-464)
//   all/any: && / || over bool        count_nonzero (and sum of bool): usize count                          (:466-715)
// Float max/min ignore NaN and start from the finite extreme (rstsr-dtype-traits/src/ext_real.rs:70-87);
// the accumulator has the element type (f32 sums in f32); integer sums wrap.
//
// A reduction is a POLICY {TI element, S state, TO output; init, pre(x, idx) -> S, comb(S, S), fin(S, n) -> TO}.
//   reduce_rows_kernel<P>  a group of G threads (1..256) owns one output and strides over the reduced index
//                          space; when the smallest-stride reduced dim is contiguous it is read as 32-byte packs
//                          (256-bit LDG), 8 packs in flight per thread.  Shuffle tree in a warp, smem tree across.
//   reduce_cols_kernel<P>  the kept fastest axis is contiguous in input and output: lanes own columns
//                          (32-byte packs), warps walk the reduced rows: coalesced >= 1 KiB per warp and row.
// Both take a split factor S along the reduced space (gridDim.y); S > 1 writes partial STATES[S][n_out] into the
// device handle's workspace and a second launch folds them in a FIXED order: results are run-to-run
// deterministic (no float atomics), which the reference's rayon fold is not.
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
    int64_t n_red;              // reduced element count (mean / var divisor)
    int group;                  // rows kernel: threads per output
    int tcol;                   // cols kernel: threads along the kept axis
    int to_partial;             // write un-finalised states to the partial buffer
    // binary reductions (vecdot, allclose): strides of the second input and the policy's run-time parameters
    int64_t ks_in2[KMAXD], rs2[KMAXD];
    double fp0, fp1;
    int ip0;
};

// A policy with `static constexpr bool BINARY = true` reads two inputs and provides pre2(x, y, desc).
template <class P, class = void> struct is_binary : std::false_type {};
template <class P> struct is_binary<P, std::void_t<decltype(P::BINARY)>> : std::integral_constant<bool, P::BINARY> {};

template <class T> using uns = typename std::make_unsigned<T>::type;

// ---------------- policies ----------------
template <class T> struct PSum {
    using TI = T; using S = T; using TO = T; using Second = PSum<T>;
    static __device__ __forceinline__ S init() { return (T)0; }
    static __device__ __forceinline__ S pre(T x, int64_t) { return x; }
    static __device__ __forceinline__ S comb(S a, S b) {
        if constexpr (std::is_integral<T>::value) return (T)((uns<T>)a + (uns<T>)b); else return a + b;
    }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};
template <class T> struct PMean {
    using TI = T; using S = T; using TO = T; using Second = PMean<T>;
    static __device__ __forceinline__ S init() { return (T)0; }
    static __device__ __forceinline__ S pre(T x, int64_t) { return x; }
    static __device__ __forceinline__ S comb(S a, S b) { return a + b; }
    static __device__ __forceinline__ TO fin(S s, int64_t n) { return s / (T)n; }  // T::from_usize(n)
};
template <class T> struct PProd {
    using TI = T; using S = T; using TO = T; using Second = PProd<T>;
    static __device__ __forceinline__ S init() { return (T)1; }
    static __device__ __forceinline__ S pre(T x, int64_t) { return x; }
    static __device__ __forceinline__ S comb(S a, S b) {
        if constexpr (std::is_integral<T>::value) return (T)((uns<T>)a * (uns<T>)b); else return a * b;
    }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};
template <class T> struct PMax {
    using TI = T; using S = T; using TO = T; using Second = PMax<T>;
    static __device__ __forceinline__ S init() { return std::numeric_limits<T>::lowest(); }
    static __device__ __forceinline__ S pre(T x, int64_t) { return x; }
    static __device__ __forceinline__ S comb(S a, S b) {
        if constexpr (std::is_same<T, float>::value) return fmaxf(a, b);
        else if constexpr (std::is_same<T, double>::value) return fmax(a, b);
        else return a < b ? b : a;
    }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};
template <class T> struct PMin {
    using TI = T; using S = T; using TO = T; using Second = PMin<T>;
    static __device__ __forceinline__ S init() { return std::numeric_limits<T>::max(); }
    static __device__ __forceinline__ S pre(T x, int64_t) { return x; }
    static __device__ __forceinline__ S comb(S a, S b) {
        if constexpr (std::is_same<T, float>::value) return fminf(a, b);
        else if constexpr (std::is_same<T, double>::value) return fmin(a, b);
        else return b < a ? b : a;
    }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};

// second pass over already-reduced states of policy P
template <class P> struct PState {
    using TI = typename P::S; using S = typename P::S; using TO = typename P::TO; using Second = PState<P>;
    static __device__ __forceinline__ S init() { return P::init(); }
    static __device__ __forceinline__ S pre(S x, int64_t) { return x; }
    static __device__ __forceinline__ S comb(S a, S b) { return P::comb(a, b); }
    static __device__ __forceinline__ TO fin(S s, int64_t n) { return P::fin(s, n); }
};

template <class T> struct PL2 {  // l2_norm: f = acc + x*x, f_out = sqrt (auto_impl/reduction.rs:319-354)
    using TI = T; using S = T; using TO = T; using Second = PState<PL2<T>>;
    static __device__ __forceinline__ S init() { return (T)0; }
    static __device__ __forceinline__ S pre(T x, int64_t) { return x * x; }
    static __device__ __forceinline__ S comb(S a, S b) { return a + b; }
    static __device__ __forceinline__ TO fin(S s, int64_t) { if constexpr (sizeof(T) == 4) return sqrtf(s); else return sqrt(s); }
};

template <class T> struct alignas(2 * sizeof(T)) Pair { T s, q; };
template <class T, bool STD> struct PVar {  // (sum, sum of squares) -> q/n - (s/n)^2 (auto_impl/reduction.rs:207-317)
    using TI = T; using S = Pair<T>; using TO = T; using Second = PState<PVar<T, STD>>;
    static __device__ __forceinline__ S init() { return S{(T)0, (T)0}; }
    static __device__ __forceinline__ S pre(T x, int64_t) { return S{x, x * x}; }
    static __device__ __forceinline__ S comb(S a, S b) { return S{a.s + b.s, a.q + b.q}; }
    static __device__ __forceinline__ TO fin(S v, int64_t n) {
        const T mean = v.s / (T)n;
        const T var = v.q / (T)n - mean * mean;
        if constexpr (!STD) return var;
        else if constexpr (sizeof(T) == 4) return sqrtf(var);
        else return sqrt(var);
    }
};

template <class T> struct PCount {  // count_nonzero / sum of bool -> usize (auto_impl/reduction.rs:522-570,678-715)
    using TI = T; using S = uint64_t; using TO = uint64_t; using Second = PState<PCount<T>>;
    static __device__ __forceinline__ S init() { return 0; }
    static __device__ __forceinline__ S pre(T x, int64_t) { return x != (T)0 ? 1 : 0; }
    static __device__ __forceinline__ S comb(S a, S b) { return a + b; }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};
template <bool ALL> struct PLogic {  // all / any over bool (auto_impl/reduction.rs:466-520)
    using TI = uint8_t; using S = uint8_t; using TO = uint8_t; using Second = PLogic<ALL>;
    static __device__ __forceinline__ S init() { return ALL ? 1 : 0; }
    static __device__ __forceinline__ S pre(uint8_t x, int64_t) { return x ? 1 : 0; }
    static __device__ __forceinline__ S comb(S a, S b) { return ALL ? (a & b) : (a | b); }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};

template <class T> struct alignas(16) ArgState { T v; int64_t i; };  // i < 0: empty ("None")
// argmin / argmax: first occurrence in row-major order wins; NaN is never accepted unless it is element 0
// (f_comp(None, y) = true, y < NaN = false: cpu_serial/reduction.rs:436-470) -- reproduced order-independently.
// The reference takes element 0 unconditionally and then only strictly better values.  Here an EMPTY state carries the
// worst value of the type (lowest for max, largest for min), so "strictly better than the accumulator" is the whole test
// in the inner loop (the first version tested emptiness, NaN and the index per element: 30 instructions per element,
// 109 registers, 2 CTAs per SM, 3.9 TB/s on an (8192, 8192) f64 argmax over rows -- ncu).  Elements equal to the worst
// value are therefore never taken; if nothing is, the result is still empty and the answer is index 0 -- exactly what the
// reference returns when no element is strictly better than element 0.  A NaN in element 0 is taken (and sticks).
template <class T, bool MAX> struct PArg {
    using TI = T; using S = ArgState<T>; using TO = uint64_t; using Second = PState<PArg<T, MAX>>;
    static __device__ __forceinline__ bool isnan_(T x) { return x != x; }
    static __device__ __forceinline__ T worst() {
        if constexpr (std::is_floating_point<T>::value) return MAX ? -INFINITY : INFINITY;
        else return MAX ? std::numeric_limits<T>::lowest() : std::numeric_limits<T>::max();
    }
    static __device__ __forceinline__ S init() { return S{worst(), -1}; }
    static __device__ __forceinline__ S pre(T x, int64_t idx) { return (isnan_(x) && idx != 0) ? init() : S{x, idx}; }
    static __device__ __forceinline__ S comb(S a, S b) {
        if (isnan_(a.v)) return a;  // only global element 0 can carry NaN: it sticks
        if (isnan_(b.v)) return b;
        const bool b_better = MAX ? (b.v > a.v) : (b.v < a.v);
        const bool a_better = MAX ? (a.v > b.v) : (a.v < b.v);
        if (b_better) return b;
        if (a_better) return a;
        return ((uint64_t)b.i < (uint64_t)a.i) ? b : a;  // tie: first occurrence; an empty state (i = -1) loses
    }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s.i < 0 ? (uint64_t)0 : (uint64_t)s.i; }
    // In-thread accumulation: every accumulator sees strictly increasing indices, so the strict comparison keeps the
    // first occurrence; it is false for a NaN element and once the accumulator holds the NaN of element 0.
    static __device__ __forceinline__ void update(S &a, T x, int64_t idx) {
        const bool better = MAX ? (x > a.v) : (x < a.v);
        if (better || (isnan_(x) && idx == 0)) { a.v = x; a.i = idx; }
    }
};

// acc <- acc (+) element: the policy's fast in-thread update when it has one, otherwise comb(acc, pre(x, idx))
template <class P, class = void> struct has_update : std::false_type {};
template <class P> struct has_update<P, std::void_t<decltype(&P::update)>> : std::true_type {};
template <class P>
__device__ __forceinline__ void fold(typename P::S &acc, typename P::TI x, int64_t idx) {
    if constexpr (has_update<P>::value) P::update(acc, x, idx);
    else acc = P::comb(acc, P::pre(x, idx));
}

// offset of linear index `i` over dims [first, n) with the given strides (cold path)
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

template <class S>
__device__ __forceinline__ S shfl_xor_state(S v, int m) {
    if constexpr (sizeof(S) < 4) {
        int x = (int)v;
        x = __shfl_xor_sync(0xffffffffu, x, m);
        return (S)x;
    } else {
        static_assert(sizeof(S) % 4 == 0, "state size");
        S r;
        const int *src = reinterpret_cast<const int *>(&v);
        int *dst = reinterpret_cast<int *>(&r);
#pragma unroll
        for (int w = 0; w < (int)(sizeof(S) / 4); ++w) dst[w] = __shfl_xor_sync(0xffffffffu, src[w], m);
        return r;
    }
}

// ------------------------------------------------------------------------------------------------
// MULTI: the reduced index space has more than one (non-mergeable) dim -> per-item decomposition.
template <class P, int VEC, bool MULTI>
__global__ void __launch_bounds__(RED_BLOCK) reduce_rows_kernel(const __grid_constant__ RedDesc d,
                                                                const typename P::TI *__restrict__ in,
                                                                const typename P::TI *__restrict__ in2,
                                                                typename P::TO *__restrict__ out,
                                                                typename P::S *__restrict__ partial) {
    using TI = typename P::TI;
    using S = typename P::S;
    constexpr bool BIN = is_binary<P>::value;
    constexpr int U = BIN ? RED_UNROLL / 2 : RED_UNROLL;  // same bytes in flight with two input streams
    __shared__ S warp_acc[RED_BLOCK / 32];
    const int G = d.group;
    const int tid = threadIdx.x;
    const int g = tid / G, t = tid - g * G;
    const int64_t o = (int64_t)blockIdx.x * (RED_BLOCK / G) + g;
    const bool valid = o < d.n_out;
    int64_t off_out = 0;
    const TI *src = in, *src2 = in2;
    if (valid) {
        src += decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_in, d.big);
        if constexpr (BIN) src2 += decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_in2, d.big);
        off_out = decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_out, d.big);
    }
    const int64_t begin = (int64_t)blockIdx.y * d.chunk;
    int64_t end = begin + d.chunk;
    if (end > d.n_items) end = d.n_items;
    if (!valid) end = begin;

    S acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = P::init();

    const int64_t rs0 = d.rs[0];  // per item (already multiplied by VEC for packs)
    const int64_t rs20 = BIN ? d.rs2[0] : 0;
    int64_t i = begin + t;
    // full groups: U independent loads per input in flight, no per-load predicate
    for (; i + (int64_t)(U - 1) * G < end; i += (int64_t)G * U) {
        Pack<TI, VEC> p[U];
        Pack<TI, VEC> q[BIN ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t ii = i + (int64_t)u * G;
            int64_t off, off2 = 0;
            if constexpr (MULTI) {
                off = decompose(ii, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
                if constexpr (BIN) off2 = decompose(ii, 0, d.nr, d.rshape, d.rdiv, d.rs2, d.big);
            } else {
                off = ii * rs0;
                off2 = ii * rs20;
            }
            p[u] = ld_stream<TI, VEC>(src + off);
            if constexpr (BIN) q[u] = ld_stream<TI, VEC>(src2 + off2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if constexpr (BIN) acc[j] = P::comb(acc[j], P::pre2(p[u].v[j], q[u].v[j], d));
                else fold<P>(acc[j], p[u].v[j], (i + (int64_t)u * G) * VEC + j);
            }
    }
    for (; i < end; i += G) {
        int64_t off, off2 = 0;
        if constexpr (MULTI) {
            off = decompose(i, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
            if constexpr (BIN) off2 = decompose(i, 0, d.nr, d.rshape, d.rdiv, d.rs2, d.big);
        } else {
            off = i * rs0;
            off2 = i * rs20;
        }
        Pack<TI, VEC> p = ld_stream<TI, VEC>(src + off);
        if constexpr (BIN) {
            Pack<TI, VEC> q = ld_stream<TI, VEC>(src2 + off2);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = P::comb(acc[j], P::pre2(p.v[j], q.v[j], d));
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) fold<P>(acc[j], p.v[j], i * VEC + j);
        }
    }
    // fold the pack lanes, then the group, in a fixed order
    S v = acc[0];
#pragma unroll
    for (int j = 1; j < VEC; ++j) v = P::comb(v, acc[j]);

    if (G <= 32) {
        for (int m = G >> 1; m >= 1; m >>= 1) v = P::comb(v, shfl_xor_state<S>(v, m));
    } else {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v = P::comb(v, shfl_xor_state<S>(v, m));
        if ((tid & 31) == 0) warp_acc[tid >> 5] = v;
        __syncthreads();
        if (t == 0) {
            const int w0 = tid >> 5, nw = G >> 5;
            v = warp_acc[w0];
            for (int w = 1; w < nw; ++w) v = P::comb(v, warp_acc[w0 + w]);
        }
    }
    if (valid && t == 0) {
        if (d.to_partial) partial[(int64_t)blockIdx.y * d.n_out_total + o] = v;
        else out[off_out] = P::fin(v, d.n_red);
    }
}

// ------------------------------------------------------------------------------------------------
// column reduction over a FEW rows (n_red <= RED_SMALL_ROWS: sums over xyz components, small batch axes).  The general
// column kernel gives such a launch 256 columns x n_red rows per CTA -- a few KB -- and pays a shared-memory fold and a
// barrier for row-lanes that have nothing to do: (3, n) f64 sum axis 0 ran at 1.9 TB/s.  Here a thread owns CPT column
// packs, walks all rows itself (CPT loads in flight per row step, 4 row steps unrolled for narrow loads) and writes the
// results: no shared memory, no barrier, 4 x the bytes per CTA.
constexpr int RED_SMALL_ROWS = 16;
template <class TI, int VEC> __host__ __device__ constexpr int red_small_cpt() { return 4; }  // 8 for narrow loads: (3, n) f64 5.4 -> 2.3 TB/s
template <class P, int VEC, bool MULTI>
__global__ void __launch_bounds__(RED_BLOCK) reduce_cols_small_kernel(const __grid_constant__ RedDesc d,
                                                                      const typename P::TI *__restrict__ in,
                                                                      const typename P::TI *__restrict__ in2,
                                                                      typename P::TO *__restrict__ out) {
    using TI = typename P::TI;
    using S = typename P::S;
    using TO = typename P::TO;
    constexpr bool BIN = is_binary<P>::value;
    constexpr int CPT = red_small_cpt<TI, VEC>();
    const int64_t total = d.packs0 * d.n_out;  // column packs over all kept dims
    const int64_t base = (int64_t)blockIdx.x * (RED_BLOCK * CPT) + threadIdx.x;
    const TI *src[CPT];
    const TI *src2[CPT];
    int64_t off_out[CPT];
    bool valid[CPT];
    S acc[CPT][VEC];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int64_t g = base + (int64_t)c * RED_BLOCK;
        valid[c] = g < total;
        src[c] = in; src2[c] = in2; off_out[c] = 0;
        if (valid[c]) {
            const int64_t kb = g / d.packs0, col = g - kb * d.packs0;
            src[c] = in + col * VEC + decompose(kb, 1, d.nk, d.kshape, d.kdiv, d.ks_in, d.big);
            if constexpr (BIN) src2[c] = in2 + col * VEC + decompose(kb, 1, d.nk, d.kshape, d.kdiv, d.ks_in2, d.big);
            off_out[c] = col * VEC + decompose(kb, 1, d.nk, d.kshape, d.kdiv, d.ks_out, d.big);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[c][j] = P::init();
    }
    const int64_t rs0 = d.rs[0], rs20 = BIN ? d.rs2[0] : 0;
    const int n = (int)d.n_items;
    constexpr int RU = VEC * sizeof(TI) >= 16 ? 1 : 4;
#pragma unroll(RU)
    for (int r = 0; r < n; ++r) {
        int64_t off, off2 = 0;
        if constexpr (MULTI) {
            off = decompose(r, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
            if constexpr (BIN) off2 = decompose(r, 0, d.nr, d.rshape, d.rdiv, d.rs2, d.big);
        } else {
            off = r * rs0;
            off2 = r * rs20;
        }
        Pack<TI, VEC> p[CPT];
        Pack<TI, VEC> q[BIN ? CPT : 1];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            if (valid[c]) {
                p[c] = ld_stream<TI, VEC>(src[c] + off);
                if constexpr (BIN) q[c] = ld_stream<TI, VEC>(src2[c] + off2);
            }
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            if (valid[c]) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if constexpr (BIN) acc[c][j] = P::comb(acc[c][j], P::pre2(p[c].v[j], q[c].v[j], d));
                    else fold<P>(acc[c][j], p[c].v[j], (int64_t)r);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        if (!valid[c]) continue;
        Pack<TO, VEC> o;
#pragma unroll
        for (int j = 0; j < VEC; ++j) o.v[j] = P::fin(acc[c][j], d.n_red);
        if constexpr (VEC > 1 && (VEC * sizeof(TO) == 16 || VEC * sizeof(TO) == 32)) {
            st_stream<TO, VEC>(out + off_out[c], o);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) out[off_out[c] + j] = o.v[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <class P, int VEC, bool MULTI>
__global__ void __launch_bounds__(RED_BLOCK) reduce_cols_kernel(const __grid_constant__ RedDesc d,
                                                                const typename P::TI *__restrict__ in,
                                                                const typename P::TI *__restrict__ in2,
                                                                typename P::TO *__restrict__ out,
                                                                typename P::S *__restrict__ partial) {
    using TI = typename P::TI;
    using S = typename P::S;
    using TO = typename P::TO;
    constexpr bool BIN = is_binary<P>::value;
    constexpr int U = BIN ? RED_UNROLL / 2 : RED_UNROLL;  // (16 for scalar loads was tried: the longer tail loses 35 %)
    extern __shared__ __align__(32) unsigned char red_smem[];
    S *sm = reinterpret_cast<S *>(red_smem);  // [RW][TC][VEC]
    const int TC = d.tcol, RW = RED_BLOCK / TC;
    const int tx = threadIdx.x % TC, ty = threadIdx.x / TC;
    const int64_t ntile0 = (d.packs0 + TC - 1) / TC;
    const int64_t tile = blockIdx.x;
    const int64_t kb = tile / ntile0;                    // linear index over kept dims 1..
    const int64_t col = (tile - kb * ntile0) * TC + tx;  // pack index along kept dim 0
    const bool valid = col < d.packs0;
    int64_t off_out = 0;
    const TI *src = in, *src2 = in2;
    if (valid) {
        src += col * VEC + decompose(kb, 1, d.nk, d.kshape, d.kdiv, d.ks_in, d.big);
        if constexpr (BIN) src2 += col * VEC + decompose(kb, 1, d.nk, d.kshape, d.kdiv, d.ks_in2, d.big);
        off_out = col * VEC + decompose(kb, 1, d.nk, d.kshape, d.kdiv, d.ks_out, d.big);
    }
    const int64_t begin = (int64_t)blockIdx.y * d.chunk;
    int64_t end = begin + d.chunk;
    if (end > d.n_items) end = d.n_items;
    if (!valid) end = begin;

    S acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = P::init();

    const int64_t rs0 = d.rs[0];
    const int64_t rs20 = BIN ? d.rs2[0] : 0;
    int64_t r = begin + ty;
    for (; r + (int64_t)(U - 1) * RW < end; r += (int64_t)RW * U) {
        Pack<TI, VEC> p[U];
        Pack<TI, VEC> q[BIN ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = r + (int64_t)u * RW;
            int64_t off, off2 = 0;
            if constexpr (MULTI) {
                off = decompose(rr, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
                if constexpr (BIN) off2 = decompose(rr, 0, d.nr, d.rshape, d.rdiv, d.rs2, d.big);
            } else {
                off = rr * rs0;
                off2 = rr * rs20;
            }
            p[u] = ld_stream<TI, VEC>(src + off);
            if constexpr (BIN) q[u] = ld_stream<TI, VEC>(src2 + off2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if constexpr (BIN) acc[j] = P::comb(acc[j], P::pre2(p[u].v[j], q[u].v[j], d));
                else fold<P>(acc[j], p[u].v[j], r + (int64_t)u * RW);
            }
    }
#pragma unroll 4
    for (; r < end; r += RW) {
        int64_t off, off2 = 0;
        if constexpr (MULTI) {
            off = decompose(r, 0, d.nr, d.rshape, d.rdiv, d.rs, d.big);
            if constexpr (BIN) off2 = decompose(r, 0, d.nr, d.rshape, d.rdiv, d.rs2, d.big);
        } else {
            off = r * rs0;
            off2 = r * rs20;
        }
        Pack<TI, VEC> p = ld_stream<TI, VEC>(src + off);
        if constexpr (BIN) {
            Pack<TI, VEC> q = ld_stream<TI, VEC>(src2 + off2);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = P::comb(acc[j], P::pre2(p.v[j], q.v[j], d));
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) fold<P>(acc[j], p.v[j], r);
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) sm[(ty * TC + tx) * VEC + j] = acc[j];
    __syncthreads();
    if (ty == 0 && valid) {
        for (int w = 1; w < RW; ++w) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = P::comb(acc[j], sm[(w * TC + tx) * VEC + j]);
        }
        if (d.to_partial) {
            S *dst = partial + (int64_t)blockIdx.y * d.n_out_total + kb * (d.packs0 * VEC) + col * VEC;
#pragma unroll
            for (int j = 0; j < VEC; ++j) dst[j] = acc[j];
        } else {
            Pack<TO, VEC> o;
#pragma unroll
            for (int j = 0; j < VEC; ++j) o.v[j] = P::fin(acc[j], d.n_red);
            if constexpr (VEC > 1 && (VEC * sizeof(TO) == 16 || VEC * sizeof(TO) == 32)) {
                st_stream<TO, VEC>(out + off_out, o);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) out[off_out + j] = o.v[j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Contiguous reduced rows that cannot be read as aligned packs from their first element (a[:, 1:-1], odd row
// lengths or pitches): each group peels the scalar head up to the first 32-byte boundary of ITS row, streams the
// aligned body as packs and finishes with the scalar tail -- no read outside the row.  One group per output, no
// split: used when there are enough rows to fill the machine.
template <class P, int VEC>
__global__ void __launch_bounds__(RED_BLOCK) reduce_rows_peel_kernel(const __grid_constant__ RedDesc d,
                                                                     const typename P::TI *__restrict__ in,
                                                                     typename P::TO *__restrict__ out) {
    using TI = typename P::TI;
    using S = typename P::S;
    __shared__ S warp_acc[RED_BLOCK / 32];
    const int G = d.group;
    const int tid = threadIdx.x;
    const int g = tid / G, t = tid - g * G;
    const int64_t o = (int64_t)blockIdx.x * (RED_BLOCK / G) + g;
    const bool valid = o < d.n_out;
    int64_t off_out = 0;
    const TI *src = in;
    int64_t n = 0;
    if (valid) {
        src += decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_in, d.big);
        off_out = decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_out, d.big);
        n = d.n_items;  // elements of the row
    }
    const int64_t mis = (int64_t)((reinterpret_cast<uintptr_t>(src) / sizeof(TI)) % VEC);
    int64_t head = mis ? VEC - mis : 0;
    if (head > n) head = n;
    const int64_t body = (n - head) / VEC;            // aligned packs
    const int64_t tail0 = head + body * VEC;          // first tail element

    S acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = P::init();
    // head / tail elements are loaded up front and folded after the body, so their latency overlaps the stream
    const bool has_h = t < head, has_t = tail0 + t < n;
    TI xh{}, xt{};
    if (has_h) xh = src[t];
    if (has_t) xt = src[tail0 + t];
    const TI *bsrc = src + head;
    int64_t i = t;
    for (; i + (int64_t)(RED_UNROLL - 1) * G < body; i += (int64_t)G * RED_UNROLL) {
        Pack<TI, VEC> p[RED_UNROLL];
#pragma unroll
        for (int u = 0; u < RED_UNROLL; ++u) p[u] = ld_stream<TI, VEC>(bsrc + (i + (int64_t)u * G) * VEC);
#pragma unroll
        for (int u = 0; u < RED_UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) fold<P>(acc[j], p[u].v[j], head + (i + (int64_t)u * G) * VEC + j);
    }
    for (; i < body; i += G) {
        Pack<TI, VEC> p = ld_stream<TI, VEC>(bsrc + i * VEC);
#pragma unroll
        for (int j = 0; j < VEC; ++j) fold<P>(acc[j], p.v[j], head + i * VEC + j);
    }
    S v = acc[0];
#pragma unroll
    for (int j = 1; j < VEC; ++j) v = P::comb(v, acc[j]);
    if (has_h) v = P::comb(v, P::pre(xh, (int64_t)t));
    if (has_t) v = P::comb(v, P::pre(xt, tail0 + t));
    if (G <= 32) {
        for (int m = G >> 1; m >= 1; m >>= 1) v = P::comb(v, shfl_xor_state<S>(v, m));
    } else {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v = P::comb(v, shfl_xor_state<S>(v, m));
        if ((tid & 31) == 0) warp_acc[tid >> 5] = v;
        __syncthreads();
        if (t == 0) {
            const int w0 = tid >> 5, nw = G >> 5;
            v = warp_acc[w0];
            for (int w = 1; w < nw; ++w) v = P::comb(v, warp_acc[w0 + w]);
        }
    }
    if (valid && t == 0) out[off_out] = P::fin(v, d.n_red);
}

// ------------------------------------------------------------------------------------------------
// Very short reduced runs (sum over the xyz axis of an (N, 3) array, (N, 4), ...): one thread per output has a single
// 32-byte load in flight (measured 2.7 TB/s for (2^24, 4) f64).  Here a thread owns TINY_OUT outputs, CTA-strided so
// that neighbouring threads read neighbouring rows, and issues the loads of all of them before folding.
constexpr int TINY_OUT = 4;
constexpr int TINY_MAX_ITEMS = 8;

template <class P, int VEC>
__global__ void __launch_bounds__(RED_BLOCK) reduce_tiny_kernel(const __grid_constant__ RedDesc d,
                                                                const typename P::TI *__restrict__ in,
                                                                const typename P::TI *__restrict__ in2,
                                                                typename P::TO *__restrict__ out) {
    using TI = typename P::TI;
    using S = typename P::S;
    constexpr bool BIN = is_binary<P>::value;
    const int64_t o0 = (int64_t)blockIdx.x * (RED_BLOCK * TINY_OUT) + threadIdx.x;
    const TI *src[TINY_OUT], *src2[TINY_OUT];
    int64_t off_out[TINY_OUT];
    bool valid[TINY_OUT];
    S acc[TINY_OUT];
#pragma unroll
    for (int u = 0; u < TINY_OUT; ++u) {
        int64_t o = o0 + (int64_t)u * RED_BLOCK;
        valid[u] = o < d.n_out;
        if (!valid[u]) o = 0;  // read a valid row, write nothing: no predicate on the loads
        src[u] = in + decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_in, d.big);
        src2[u] = in2;
        if constexpr (BIN) src2[u] = in2 + decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_in2, d.big);
        off_out[u] = decompose(o, 0, d.nk, d.kshape, d.kdiv, d.ks_out, d.big);
        acc[u] = P::init();
    }
    const int n = (int)d.n_items;
    const int64_t rs0 = d.rs[0], rs20 = BIN ? d.rs2[0] : 0;
    for (int i = 0; i < n; ++i) {
        Pack<TI, VEC> p[TINY_OUT];
        Pack<TI, VEC> q[BIN ? TINY_OUT : 1];
#pragma unroll
        for (int u = 0; u < TINY_OUT; ++u) {
            p[u] = ld_stream<TI, VEC>(src[u] + i * rs0);
            if constexpr (BIN) q[u] = ld_stream<TI, VEC>(src2[u] + i * rs20);
        }
#pragma unroll
        for (int u = 0; u < TINY_OUT; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if constexpr (BIN) acc[u] = P::comb(acc[u], P::pre2(p[u].v[j], q[u].v[j], d));
                else fold<P>(acc[u], p[u].v[j], (int64_t)i * VEC + j);
            }
    }
#pragma unroll
    for (int u = 0; u < TINY_OUT; ++u)
        if (valid[u]) out[off_out[u]] = P::fin(acc[u], d.n_red);
}

template <class P, int V>
void launch_tiny(rc_device *dev, const RedDesc &d, int vec, const typename P::TI *in, const typename P::TI *in2,
                 typename P::TO *out) {
    const int64_t gx = (d.n_out + RED_BLOCK * TINY_OUT - 1) / (RED_BLOCK * TINY_OUT);
    RC_CHECK(gx < (1ll << 31), RC_ERR_UNIMPLEMENTED, "reduction grid too large");
    if constexpr (V > 1) {
        if (vec > 1) {
            reduce_tiny_kernel<P, V><<<(unsigned)gx, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out);
            after_launch(dev, "reduce_tiny_kernel");
            return;
        }
    }
    reduce_tiny_kernel<P, 1><<<(unsigned)gx, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out);
    after_launch(dev, "reduce_tiny_kernel");
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
        if (c.binary) d.ks_in2[i] = c.kstride_in2[i];
        if (c.kshape[i] >= (1ll << 31)) d.big = 1; else d.kdiv[i] = FastDiv((uint32_t)c.kshape[i]);
    }
    for (int i = 0; i < d.nr; ++i) {
        d.rshape[i] = c.rshape[i];
        d.rs[i] = c.rstride[i];
        if (c.binary) d.rs2[i] = c.rstride2[i];
        if (c.rshape[i] >= (1ll << 31)) d.big = 1; else d.rdiv[i] = FastDiv((uint32_t)std::max<int64_t>(c.rshape[i], 1));
    }
}

inline bool aligned_bytes(const void *p, size_t bytes) { return reinterpret_cast<uintptr_t>(p) % bytes == 0; }

template <class P, int V>
void launch_rows(rc_device *dev, const RedDesc &d, int vec, int64_t sy, const typename P::TI *in,
                 const typename P::TI *in2, typename P::TO *out, typename P::S *partial) {
    int64_t gx = (d.n_out + (RED_BLOCK / d.group) - 1) / (RED_BLOCK / d.group);
    RC_CHECK(gx < (1ll << 31) && sy <= 65535, RC_ERR_UNIMPLEMENTED, "reduction grid too large");
    dim3 grid((unsigned)gx, (unsigned)sy);
    const bool multi = d.nr > 1;
    if constexpr (V > 1) {
        if (vec > 1) {
            if (multi) reduce_rows_kernel<P, V, true><<<grid, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out, partial);
            else reduce_rows_kernel<P, V, false><<<grid, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out, partial);
            after_launch(dev, "reduce_rows_kernel");
            return;
        }
    }
    if (multi) reduce_rows_kernel<P, 1, true><<<grid, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out, partial);
    else reduce_rows_kernel<P, 1, false><<<grid, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out, partial);
    after_launch(dev, "reduce_rows_kernel");
}

template <class P, int V>
void launch_cols(rc_device *dev, const RedDesc &d, int vec, int64_t sy, const typename P::TI *in,
                 const typename P::TI *in2, typename P::TO *out, typename P::S *partial) {
    int64_t ntile0 = (d.packs0 + d.tcol - 1) / d.tcol;
    int64_t gx = ntile0 * d.n_out;
    RC_CHECK(gx < (1ll << 31) && sy <= 65535, RC_ERR_UNIMPLEMENTED, "reduction grid too large");
    dim3 grid((unsigned)gx, (unsigned)sy);
    size_t smem = (size_t)RED_BLOCK * sizeof(typename P::S) * (vec > 1 ? V : 1);
    const bool multi = d.nr > 1;
    if constexpr (V > 1) {
        if (vec > 1) {
            if (smem > 48 * 1024) {  // wide packs of a narrow element with a wide state (bool -> u64 counts)
                RC_CUDA(cudaFuncSetAttribute(reduce_cols_kernel<P, V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                RC_CUDA(cudaFuncSetAttribute(reduce_cols_kernel<P, V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
            if (multi) reduce_cols_kernel<P, V, true><<<grid, RED_BLOCK, smem, dev->stream>>>(d, in, in2, out, partial);
            else reduce_cols_kernel<P, V, false><<<grid, RED_BLOCK, smem, dev->stream>>>(d, in, in2, out, partial);
            after_launch(dev, "reduce_cols_kernel");
            return;
        }
    }
    if (multi) reduce_cols_kernel<P, 1, true><<<grid, RED_BLOCK, smem, dev->stream>>>(d, in, in2, out, partial);
    else reduce_cols_kernel<P, 1, false><<<grid, RED_BLOCK, smem, dev->stream>>>(d, in, in2, out, partial);
    after_launch(dev, "reduce_cols_kernel");
}

template <class P, int V>
bool launch_cols_small(rc_device *dev, const RedDesc &d, int vec, const typename P::TI *in, const typename P::TI *in2,
                       typename P::TO *out) {
    static const bool off = [] { const char *e = getenv("RC_COLS_SMALL"); return e && e[0] == '0'; }();
    using TI = typename P::TI;
    const int64_t total = d.packs0 * d.n_out;
    const int64_t per_cta = (int64_t)RED_BLOCK * (vec > 1 ? red_small_cpt<TI, V>() : red_small_cpt<TI, 1>());
    const int64_t gx = (total + per_cta - 1) / per_cta;
    if (off || gx >= (1ll << 31)) return false;
    const bool multi = d.nr > 1;
    if constexpr (V > 1) {
        if (vec > 1) {
            if (multi) reduce_cols_small_kernel<P, V, true><<<(unsigned)gx, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out);
            else reduce_cols_small_kernel<P, V, false><<<(unsigned)gx, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out);
            after_launch(dev, "reduce_cols_small_kernel");
            return true;
        }
    }
    if (multi) reduce_cols_small_kernel<P, 1, true><<<(unsigned)gx, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out);
    else reduce_cols_small_kernel<P, 1, false><<<(unsigned)gx, RED_BLOCK, 0, dev->stream>>>(d, in, in2, out);
    after_launch(dev, "reduce_cols_small_kernel");
    return true;
}

// threads along the kept (contiguous) axis of the column kernel; RC_TCOL_MAX is a tuning knob for experiments
inline int tcol_max() {
    static int v = [] { const char *e = getenv("RC_TCOL_MAX"); int x = e ? atoi(e) : 64; return (x >= 1 && x <= RED_BLOCK) ? x : 64; }();
    return v;
}

// smallest contiguous kept extent that takes the column kernel; RC_COLS_MIN is a tuning knob for experiments
inline int cols_min_extent() {
    static int v = [] { const char *e = getenv("RC_COLS_MIN"); int x = e ? atoi(e) : 2; return x >= 1 ? x : 2; }();
    return v;
}

int pow2_floor(int64_t x) { int p = 1; while ((int64_t)p * 2 <= x) p *= 2; return p; }
int pow2_ceil(int64_t x) { int p = 1; while (p < x) p *= 2; return p; }

// P: first-pass policy.  The second pass (over partial states) uses P::Second (P itself when pre is the identity).
// V = elements per pack of the vector variants this instantiation may launch (it falls back to one element per load
// when extents / strides / pointers rule the packs out); see reduce_typed below for how V is chosen
template <class P, int V>
void reduce_typed_w(rc_device *dev, const CanonRed &c, const void *a_v, void *out_v, int64_t n_red_logical,
                    const void *b_v = nullptr, double fp0 = 0.0, double fp1 = 0.0, int ip0 = 0) {
    using TI = typename P::TI;
    using S = typename P::S;
    using TO = typename P::TO;
    using P2 = typename P::Second;
    if (c.empty_out) return;
    constexpr bool BIN = is_binary<P>::value;
    RC_CHECK(BIN == c.binary, RC_ERR_INVALID_VALUE, "internal: reduction arity mismatch");
    const TI *in = static_cast<const TI *>(a_v) + c.base_in;
    const TI *in2 = BIN ? static_cast<const TI *>(b_v) + c.base_in2 : nullptr;
    TO *out = static_cast<TO *>(out_v) + c.base_out;
    constexpr bool SIMPLE = std::is_same<P2, P>::value;  // state == element, pre == identity
    const int64_t n_out = c.n_out(), n_red = c.n_red();
    // two FULL waves at the kernels' occupancy (4 CTAs of 256 threads per SM at <= 64 registers): a split that
    // overshoots a wave boundary leaves a tail wave (measured: 1280 vs 1184 CTAs costs 4 %)
    const int64_t target_ctas = (int64_t)dev->sm_count * 8;

    RedDesc d;
    std::memset(&d, 0, sizeof(d));
    fill_desc_dims(d, c);
    d.n_red = n_red_logical;
    d.fp0 = fp0; d.fp1 = fp1; d.ip0 = ip0;
    d.n_out_total = n_out;
    if (n_out >= (1ll << 31)) d.big = 1;

    const bool red_contig = d.nr >= 1 && d.rs[0] == 1 && (!BIN || d.rs2[0] == 1) && d.rshape[0] >= 8;
    const bool kept_contig = d.nk >= 1 && d.ks_in[0] == 1 && (!BIN || d.ks_in2[0] == 1) && d.ks_out[0] == 1 &&
                             d.kshape[0] >= cols_min_extent();

    // a sharded reduction folds the partial states itself (fused with the cross-GPU combine): possible when the state
    // is the element (sum / prod / max / min) and output index o lives at out[o] (canonical order is contiguous)
    auto fuse_second_pass = [&]() {
        if (!dev->preq.want || !SIMPLE || BIN || !std::is_same<S, TO>::value) return false;
        int64_t acc = 1;
        for (size_t i = 0; i < c.kshape.size(); ++i) {
            if (c.kstride_out[i] != acc) return false;
            acc *= c.kshape[i];
        }
        return true;
    };
    auto second_pass_desc = [&](const RedDesc &first, int64_t Sx) {
        RedDesc e = first;
        e.to_partial = 0;
        e.nr = 1;
        e.rshape[0] = Sx;
        e.rs[0] = n_out;
        e.rdiv[0] = FastDiv((uint32_t)Sx);
        int64_t acc = 1;
        for (int i = 0; i < e.nk; ++i) { e.ks_in[i] = acc; acc *= e.kshape[i]; }
        e.n_items = Sx;
        e.chunk = Sx;
        return e;
    };

    if (!red_contig && kept_contig && n_red > 0) {
        // ---------------- column kernel ----------------
        bool vec_ok = V > 1 && d.kshape[0] % V == 0 && aligned_bytes(in, V * sizeof(TI)) && aligned_bytes(out, V * sizeof(TO));
        for (int i = 1; i < d.nk && vec_ok; ++i) vec_ok = d.ks_in[i] % V == 0 && d.ks_out[i] % V == 0;
        for (int i = 0; i < d.nr && vec_ok; ++i) vec_ok = d.rs[i] % V == 0;
        if constexpr (BIN) {
            vec_ok = vec_ok && aligned_bytes(in2, V * sizeof(TI));
            for (int i = 1; i < d.nk && vec_ok; ++i) vec_ok = d.ks_in2[i] % V == 0;
            for (int i = 0; i < d.nr && vec_ok; ++i) vec_ok = d.rs2[i] % V == 0;
        }
        const int vec = vec_ok ? V : 1;
        d.packs0 = d.kshape[0] / vec;
        // threads along the kept axis: 64 (4 warp-rows walk the reduced rows) unless the reduced extent is too short to
        // keep RED_UNROLL loads per thread in flight -- then fewer row-lanes, down to one thread per column pack
        // ((4, 2^24) sum axis 0: 2.3 -> TB/s measured with the fixed 64)
        d.n_out = n_out / d.kshape[0];
        d.n_items = n_red;
        if (n_red >= (1ll << 31) || d.n_out >= (1ll << 31)) d.big = 1;
        // a few rows: one thread per column pack, no row-lanes (reduce_cols_small_kernel).  Up to 4 rows always (2 rows
        // +29 %, 3 rows 2.8x); 5..16 rows only when the kept rows are short, where the general kernel's 64-column tiles
        // do not fill ((32768,16,256) f32 +62 %; (8, 2^24) and (256,16,32768) are 4 % better on the general kernel)
        if (n_red <= 4 || (n_red <= RED_SMALL_ROWS && d.packs0 <= 128)) {
            d.to_partial = 0;
            if (launch_cols_small<P, V>(dev, d, vec, in, in2, out)) return;
        }
        const int64_t unroll = RED_UNROLL;
        // one element per load (extents / pointers rule the packs out): wider column tiles, so that a CTA still reads
        // 512 B - 1 KB of every row it visits (sweep, scripts/probe_reduce_sweep.py with RC_TCOL_MAX: f32 (100, 1342177)
        // axis 0 3.57 / 4.07 / 4.40 TB/s at 64 / 128 / 256 columns, f64 5.79 / 6.14 / 6.26; long reduced extents prefer 128)
        const int tcol_limit = vec > 1 ? tcol_max() : std::max(tcol_max(), (sizeof(TI) <= 4 && n_red <= 512) ? 256 : 128);
        const int64_t rw_want = std::max<int64_t>(1, std::min<int64_t>(RED_BLOCK / tcol_limit, pow2_floor(std::max<int64_t>(1, n_red / unroll))));
        d.tcol = (int)std::min<int64_t>(RED_BLOCK / rw_want, pow2_ceil(d.packs0));
        d.n_out = n_out / d.kshape[0];
        d.n_items = n_red;
        if (n_red >= (1ll << 31) || d.n_out >= (1ll << 31)) d.big = 1;
        const int rw = RED_BLOCK / d.tcol;
        int64_t base_ctas = ((d.packs0 + d.tcol - 1) / d.tcol) * d.n_out;
        int64_t Sx = std::min<int64_t>(std::max<int64_t>(1, target_ctas / base_ctas),
                                       std::max<int64_t>(1, n_red / ((int64_t)rw * unroll * 2)));
        Sx = std::max<int64_t>(1, std::min<int64_t>(Sx, 1024));
        d.chunk = (n_red + Sx - 1) / Sx;
        Sx = (n_red + d.chunk - 1) / d.chunk;
        if (Sx == 1) {
            d.to_partial = 0;
            launch_cols<P, V>(dev, d, vec, 1, in, in2, out, (S *)nullptr);
            return;
        }
        S *partial = static_cast<S *>(workspace(dev, (size_t)Sx * n_out * sizeof(S)));
        d.to_partial = 1;
        launch_cols<P, V>(dev, d, vec, Sx, in, in2, out, partial);
        if (fuse_second_pass()) { dev->preq.got = true; dev->preq.ptr = partial; dev->preq.S = Sx; dev->preq.pitch = n_out; dev->preq.out = out; return; }
        // second pass: fold partial[S][n_out] over S, same kept dims with contiguous input strides
        RedDesc e = second_pass_desc(d, Sx);
        constexpr int V2 = SIMPLE ? V : 1;
        const bool vec2 = V2 > 1 && vec_ok && (n_out % V2 == 0) && aligned_bytes(partial, V2 * sizeof(S));
        const int v2 = vec2 ? V2 : 1;
        e.packs0 = e.kshape[0] / v2;
        e.tcol = (int)std::min<int64_t>(tcol_max(), pow2_ceil(e.packs0));
        launch_cols<P2, V2>(dev, e, v2, 1, partial, (const S *)nullptr, out, (S *)nullptr);
        return;
    }

    // ---------------- row kernel (contiguous or generic reduced space) ----------------
    bool vec_ok = V > 1 && d.nr >= 1 && d.rs[0] == 1 && d.rshape[0] % V == 0 && aligned_bytes(in, V * sizeof(TI));
    for (int i = 0; i < d.nk && vec_ok; ++i) vec_ok = d.ks_in[i] % V == 0;
    for (int i = 1; i < d.nr && vec_ok; ++i) vec_ok = d.rs[i] % V == 0;
    if constexpr (BIN) {
        vec_ok = vec_ok && d.rs2[0] == 1 && aligned_bytes(in2, V * sizeof(TI));
        for (int i = 0; i < d.nk && vec_ok; ++i) vec_ok = d.ks_in2[i] % V == 0;
        for (int i = 1; i < d.nr && vec_ok; ++i) vec_ok = d.rs2[i] % V == 0;
    }
    if constexpr (!BIN && V > 1 && sizeof(TI) <= 4) {
        // contiguous rows that only alignment / divisibility keeps from the pack path, and enough of them: peel per row.
        // (8-byte elements stay on the scalar path: measured 5.3 TB/s there vs 4.5-5.0 with the peel, whose 68-72
        // registers cost a CTA per SM; f32 gets 5.3 TB/s with it, 1.4x torch.)
        if (!vec_ok && d.nr == 1 && d.rs[0] == 1 && d.rshape[0] >= 16 * V && d.rshape[0] < (1ll << 31)) {
            RedDesc e = d;
            e.n_out = n_out;
            e.n_items = n_red;
            e.to_partial = 0;
            e.group = (int)std::min<int64_t>(RED_BLOCK, std::max<int64_t>(1, pow2_floor(std::max<int64_t>(1, n_red / V / RED_UNROLL))));
            const int64_t ctas = (n_out + (RED_BLOCK / e.group) - 1) / (RED_BLOCK / e.group);
            if (ctas >= 2 * (int64_t)dev->sm_count && ctas < (1ll << 31)) {
                reduce_rows_peel_kernel<P, V><<<(unsigned)ctas, RED_BLOCK, 0, dev->stream>>>(e, in, out);
                after_launch(dev, "reduce_rows_peel_kernel");
                return;
            }
        }
    }
    const int vec = vec_ok ? V : 1;
    if (d.nr == 0) {  // nothing reduced (all reduced axes have extent 1): a strided copy through the monoid
        d.nr = 1;
        d.rshape[0] = 1;
        d.rs[0] = 0;
        d.rs2[0] = 0;
        d.rdiv[0] = FastDiv(1);
    }
    if (vec > 1) {
        d.rshape[0] /= vec;
        d.rdiv[0] = FastDiv((uint32_t)std::max<int64_t>(d.rshape[0], 1));
        d.rs[0] = vec;
        d.rs2[0] = vec;
    }
    d.n_out = n_out;
    d.n_items = n_red / vec;
    if (d.n_items >= (1ll << 31) && d.nr > 1) d.big = 1;
    // many outputs, runs of at most 64 bytes ((2^22, 16) f64 is already better off with the row kernel: 5.6 vs 5.2 TB/s)
    if (d.nr == 1 && d.n_items <= TINY_MAX_ITEMS && d.n_items * vec * (int64_t)sizeof(TI) <= 64 && n_out >= 4096) {
        d.to_partial = 0;
        launch_tiny<P, V>(dev, d, vec, in, in2, out);
        return;
    }
    d.group = (int)std::min<int64_t>(RED_BLOCK, std::max<int64_t>(1, pow2_floor(std::max<int64_t>(1, d.n_items / RED_UNROLL))));
    int64_t base_ctas = (n_out + (RED_BLOCK / d.group) - 1) / (RED_BLOCK / d.group);
    int64_t Sx = 1;
    if (base_ctas < target_ctas) {
        Sx = std::min<int64_t>(std::max<int64_t>(1, target_ctas / base_ctas),
                               std::max<int64_t>(1, d.n_items / ((int64_t)d.group * RED_UNROLL * 4)));
        Sx = std::max<int64_t>(1, std::min<int64_t>(Sx, 4096));
    }
    d.chunk = (d.n_items + Sx - 1) / std::max<int64_t>(Sx, 1);
    if (d.chunk == 0) d.chunk = 1;
    Sx = std::max<int64_t>(1, (d.n_items + d.chunk - 1) / d.chunk);
    if (Sx == 1) {
        d.to_partial = 0;
        launch_rows<P, V>(dev, d, vec, 1, in, in2, out, (S *)nullptr);
        return;
    }
    S *partial = static_cast<S *>(workspace(dev, (size_t)Sx * n_out * sizeof(S)));
    d.to_partial = 1;
    launch_rows<P, V>(dev, d, vec, Sx, in, in2, out, partial);
    if (fuse_second_pass()) { dev->preq.got = true; dev->preq.ptr = partial; dev->preq.S = Sx; dev->preq.pitch = n_out; dev->preq.out = out; return; }
    // second pass: out[o] = fold_s partial[s][o]
    RedDesc e = second_pass_desc(d, Sx);
    e.group = (int)std::min<int64_t>(RED_BLOCK, std::max<int64_t>(1, pow2_floor(std::max<int64_t>(1, Sx / 2))));
    launch_rows<P2, 1>(dev, e, 1, 1, partial, (const S *)nullptr, out, (S *)nullptr);
}

// the five monoids of the hot path
// Policies that also get HALF-width packs (16 bytes): sums / extrema / means of f32 and f64, where extents that are a
// multiple of 4 but not of 8 elements (f32: 12, 20, 100 ...) or of 2 but not of 4 (f64) are common and the one-element path
// costs a factor of 2-3 ((20971, 64, 100) f32 sum over axes (0, 2): 1.7 TB/s on the scalar path, 5.7 for f64 whose packs
// fit).  Kept to these policies: every extra width is a full set of kernels per policy at build time.
template <class P> struct half_packs : std::false_type {};
template <> struct half_packs<PSum<float>> : std::true_type {};
template <> struct half_packs<PSum<double>> : std::true_type {};
template <> struct half_packs<PMax<float>> : std::true_type {};
template <> struct half_packs<PMax<double>> : std::true_type {};
template <> struct half_packs<PMin<float>> : std::true_type {};
template <> struct half_packs<PMin<double>> : std::true_type {};
template <> struct half_packs<PMean<float>> : std::true_type {};
template <> struct half_packs<PMean<double>> : std::true_type {};

// would packs of w elements pass the checks of reduce_typed_w?  (A wrong "yes" only costs the scalar fallback there.)
template <class P>
bool packs_fit(const CanonRed &c, const void *a_v, const void *b_v, const void *out_v, int w) {
    using TI = typename P::TI;
    using TO = typename P::TO;
    constexpr bool BIN = is_binary<P>::value;
    RedDesc d;
    std::memset(&d, 0, sizeof(d));
    fill_desc_dims(d, c);
    const TI *in = static_cast<const TI *>(a_v) + c.base_in;
    const TI *in2 = BIN ? static_cast<const TI *>(b_v) + c.base_in2 : nullptr;
    const TO *out = static_cast<const TO *>(out_v) + c.base_out;
    const bool red_contig = d.nr >= 1 && d.rs[0] == 1 && (!BIN || d.rs2[0] == 1) && d.rshape[0] >= 8;
    const bool kept_contig = d.nk >= 1 && d.ks_in[0] == 1 && (!BIN || d.ks_in2[0] == 1) && d.ks_out[0] == 1 &&
                             d.kshape[0] >= cols_min_extent();
    bool ok = aligned_bytes(in, w * sizeof(TI)) && (!BIN || aligned_bytes(in2, w * sizeof(TI)));
    if (!red_contig && kept_contig) {
        ok = ok && d.kshape[0] % w == 0 && aligned_bytes(out, w * sizeof(TO));
        for (int i = 1; i < d.nk && ok; ++i) ok = d.ks_in[i] % w == 0 && d.ks_out[i] % w == 0 && (!BIN || d.ks_in2[i] % w == 0);
        for (int i = 0; i < d.nr && ok; ++i) ok = d.rs[i] % w == 0 && (!BIN || d.rs2[i] % w == 0);
    } else {
        ok = ok && d.nr >= 1 && d.rs[0] == 1 && d.rshape[0] % w == 0 && (!BIN || d.rs2[0] == 1);
        for (int i = 0; i < d.nk && ok; ++i) ok = d.ks_in[i] % w == 0 && (!BIN || d.ks_in2[i] % w == 0);
        for (int i = 1; i < d.nr && ok; ++i) ok = d.rs[i] % w == 0 && (!BIN || d.rs2[i] % w == 0);
    }
    return ok;
}

template <class P>
void reduce_typed(rc_device *dev, const CanonRed &c, const void *a_v, void *out_v, int64_t n_red_logical,
                  const void *b_v = nullptr, double fp0 = 0.0, double fp1 = 0.0, int ip0 = 0) {
    constexpr int V = 32 / sizeof(typename P::TI);  // 256-bit packs of the element type
    if constexpr (half_packs<P>::value && V >= 4) {
        static const bool off = [] { const char *e = getenv("RC_RED_HALF_PACKS"); return e && e[0] == '0'; }();
        if (!off && !c.empty_out && !packs_fit<P>(c, a_v, b_v, out_v, V) && packs_fit<P>(c, a_v, b_v, out_v, V / 2)) {
            reduce_typed_w<P, V / 2>(dev, c, a_v, out_v, n_red_logical, b_v, fp0, fp1, ip0);
            return;
        }
    }
    reduce_typed_w<P, V>(dev, c, a_v, out_v, n_red_logical, b_v, fp0, fp1, ip0);
}

template <class T>
void reduce_op(rc_device *dev, rc_redop op, const CanonRed &c, const void *a, void *out, int64_t n_red) {
    switch (op) {
        case RC_SUM: reduce_typed<PSum<T>>(dev, c, a, out, n_red); return;
        case RC_PROD: reduce_typed<PProd<T>>(dev, c, a, out, n_red); return;
        case RC_MAX: reduce_typed<PMax<T>>(dev, c, a, out, n_red); return;
        case RC_MIN: reduce_typed<PMin<T>>(dev, c, a, out, n_red); return;
        case RC_MEAN:
            if constexpr (std::is_floating_point<T>::value) {
                reduce_typed<PMean<T>>(dev, c, a, out, n_red);
                return;
            } else {
                raise(RC_ERR_UNIMPLEMENTED, "mean is only defined for floating-point element types");
            }
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "unknown reduction op");
}

// the "next" reductions: var / std / l2_norm (floats), argmin / argmax, count_nonzero
template <class T>
void reduce_op_ext(rc_device *dev, rc_redop op, const CanonRed &c, const void *a, void *out, int64_t n_red) {
    switch (op) {
        case RC_ARGMIN: reduce_typed<PArg<T, false>>(dev, c, a, out, n_red); return;
        case RC_ARGMAX: reduce_typed<PArg<T, true>>(dev, c, a, out, n_red); return;
        case RC_COUNT_NONZERO: reduce_typed<PCount<T>>(dev, c, a, out, n_red); return;
        default: break;
    }
    if constexpr (std::is_floating_point<T>::value) {
        switch (op) {
            case RC_VAR: reduce_typed<PVar<T, false>>(dev, c, a, out, n_red); return;
            case RC_STD: reduce_typed<PVar<T, true>>(dev, c, a, out, n_red); return;
            case RC_L2_NORM: reduce_typed<PL2<T>>(dev, c, a, out, n_red); return;
            default: break;
        }
    }
    raise(RC_ERR_UNIMPLEMENTED, "this reduction is not implemented for the element type");
}

}  // namespace
}  // namespace rc
