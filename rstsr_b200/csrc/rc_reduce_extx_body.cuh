// rc_reduce_extx_body.cuh -- reductions of the extended element types (f16, bf16, c32, c64) on the kernels of rc_reduce.cuh.
//   half     every policy of the f32 path through PViaF32: inputs converted to f32, f32 state, ONE rounding of the result
//            (sum / prod / max / min / mean / var / std / l2_norm), index / count outputs unchanged (argmin / argmax /
//            count_nonzero).  The reference accumulates in the element type (`acc + x` in half precision): ours is the
//            more accurate value and lies within half-precision rounding of it.
//   complex  sum / prod / mean componentwise resp. with the operators of rc_types.cuh (the mean divides by
//            Complex::from(n)); var / std / l2_norm with REAL output as auto_impl/reduction.rs:207-354 defines them:
//            state (sum x, sum (x * conj x).re), var = q / n - (m * conj m).re with m = s / Complex::from(n).
// Compiled once per element type (rc_reduce_ext_{h16,b16,c32,c64}.cu define RC_EXTX_KIND and include this body): one
// translation unit with all four took over five minutes and was the long pole of the build.
#include "rc_reduce.cuh"
#include "rc_types.cuh"

namespace rc {
namespace {


// half input through a policy PF of the f32 path; outputs of type float become the half type again
template <class T, class PF>
struct PViaF32 {
    using TI = T;
    using S = typename PF::S;
    using TO = typename std::conditional<std::is_same<typename PF::TO, float>::value, T, typename PF::TO>::type;
    using Second = PState<PViaF32<T, PF>>;
    static __device__ __forceinline__ S init() { return PF::init(); }
    static __device__ __forceinline__ S pre(T x, int64_t idx) { return PF::pre(x.f(), idx); }
    static __device__ __forceinline__ S comb(S a, S b) { return PF::comb(a, b); }
    static __device__ __forceinline__ TO fin(S s, int64_t n) {
        if constexpr (std::is_same<typename PF::TO, float>::value) return T(PF::fin(s, n));
        else return PF::fin(s, n);
    }
};

template <class R> struct alignas(4 * sizeof(R)) CVarState { cplx<R> s; R q; R pad; };  // 16 / 32 bytes: one LDG
template <class R, bool STD> struct PCVar {
    using TI = cplx<R>; using S = CVarState<R>; using TO = R; using Second = PState<PCVar<R, STD>>;
    static __device__ __forceinline__ S init() { return S{cplx<R>((R)0, (R)0), (R)0, (R)0}; }
    static __device__ __forceinline__ S pre(TI x, int64_t) { return S{x, x.re * x.re + x.im * x.im, (R)0}; }
    static __device__ __forceinline__ S comb(S a, S b) { return S{a.s + b.s, a.q + b.q, (R)0}; }
    static __device__ __forceinline__ TO fin(S v, int64_t n) {
        const cplx<R> mean = v.s / cplx<R>((R)n, (R)0);
        const R var = v.q / (R)n - (mean.re * mean.re + mean.im * mean.im);
        if constexpr (!STD) return var;
        else if constexpr (sizeof(R) == 4) return sqrtf(var);
        else return sqrt(var);
    }
};
template <class R> struct PCL2 {
    using TI = cplx<R>; using S = R; using TO = R; using Second = PState<PCL2<R>>;
    static __device__ __forceinline__ S init() { return (R)0; }
    static __device__ __forceinline__ S pre(TI x, int64_t) { return x.re * x.re + x.im * x.im; }
    static __device__ __forceinline__ S comb(S a, S b) { return a + b; }
    static __device__ __forceinline__ TO fin(S s, int64_t) { if constexpr (sizeof(R) == 4) return sqrtf(s); else return sqrt(s); }
};

template <class T>
void reduce_half(rc_device *dev, rc_redop op, const CanonRed &c, const void *a, void *out, int64_t n) {
    switch (op) {
        case RC_SUM: reduce_typed<PViaF32<T, PSum<float>>>(dev, c, a, out, n); return;
        case RC_PROD: reduce_typed<PViaF32<T, PProd<float>>>(dev, c, a, out, n); return;
        case RC_MAX: reduce_typed<PViaF32<T, PMax<float>>>(dev, c, a, out, n); return;
        case RC_MIN: reduce_typed<PViaF32<T, PMin<float>>>(dev, c, a, out, n); return;
        case RC_MEAN: reduce_typed<PViaF32<T, PMean<float>>>(dev, c, a, out, n); return;
        case RC_VAR: reduce_typed<PViaF32<T, PVar<float, false>>>(dev, c, a, out, n); return;
        case RC_STD: reduce_typed<PViaF32<T, PVar<float, true>>>(dev, c, a, out, n); return;
        case RC_L2_NORM: reduce_typed<PViaF32<T, PL2<float>>>(dev, c, a, out, n); return;
        case RC_ARGMIN: reduce_typed<PViaF32<T, PArg<float, false>>>(dev, c, a, out, n); return;
        case RC_ARGMAX: reduce_typed<PViaF32<T, PArg<float, true>>>(dev, c, a, out, n); return;
        case RC_COUNT_NONZERO: reduce_typed<PViaF32<T, PCount<float>>>(dev, c, a, out, n); return;
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, "this reduction is not implemented for half types");
}

template <class R>
void reduce_cplx(rc_device *dev, rc_redop op, const CanonRed &c, const void *a, void *out, int64_t n) {
    using T = cplx<R>;
    switch (op) {
        case RC_SUM: reduce_typed<PSum<T>>(dev, c, a, out, n); return;
        case RC_PROD: reduce_typed<PProd<T>>(dev, c, a, out, n); return;
        case RC_MEAN: reduce_typed<PMean<T>>(dev, c, a, out, n); return;
        case RC_VAR: reduce_typed<PCVar<R, false>>(dev, c, a, out, n); return;
        case RC_STD: reduce_typed<PCVar<R, true>>(dev, c, a, out, n); return;
        case RC_L2_NORM: reduce_typed<PCL2<R>>(dev, c, a, out, n); return;
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, "complex numbers have no ordering: max / min / argmin / argmax are not defined for them "
                                "(ExtReal is not implemented for Complex)");
}

// ---- vecdot: sum conj(a) * b (cpu_serial/vecdot.rs:96-157; ExtNum::ext_conj is the identity on real types) ----
template <class R> struct PCDot {
    static constexpr bool BINARY = true;
    using TI = cplx<R>; using S = cplx<R>; using TO = cplx<R>; using Second = PSum<cplx<R>>;
    static __device__ __forceinline__ S init() { return cplx<R>((R)0, (R)0); }
    static __device__ __forceinline__ S pre2(TI x, TI y, const RedDesc &) { return cplx<R>(x.re, -x.im) * y; }
    static __device__ __forceinline__ S comb(S a, S b) { return a + b; }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};
// half: the product is rounded to the element type as `x * y` is, accumulated in f32, ONE rounding of the sum
template <class T> struct PHDot {
    static constexpr bool BINARY = true;
    using TI = T; using S = float; using TO = T; using Second = PState<PHDot<T>>;
    static __device__ __forceinline__ S init() { return 0.0f; }
    static __device__ __forceinline__ S pre2(T x, T y, const RedDesc &) { return (x * y).f(); }
    static __device__ __forceinline__ S comb(S a, S b) { return a + b; }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return T(s); }
};
// ---- allclose_all: |a - b| <= atol + rtol * |b| with TE = f64 (rstsr-dtype-traits/src/isclose.rs:92-106); the difference
// and |b| are formed in the element type (complex: Complex::norm = hypot) and then widened ----
template <class T> struct PCloseX {
    static constexpr bool BINARY = true;
    using TI = T; using S = uint8_t; using TO = uint8_t; using Second = PLogic<true>;
    static __device__ __forceinline__ S init() { return 1; }
    static __device__ __forceinline__ S pre2(T a, T b, const RedDesc &d) {
        double diff, abs_b;
        bool both_nan;
        if constexpr (is_cplx_t<T>::value) {
            using R = typename real_of<T>::type;
            const T df = a - b;
            if constexpr (sizeof(R) == 4) { diff = (double)hypotf(df.re, df.im); abs_b = (double)hypotf(b.re, b.im); }
            else { diff = hypot(df.re, df.im); abs_b = hypot(b.re, b.im); }
            both_nan = (a.re != a.re || a.im != a.im) && (b.re != b.re || b.im != b.im);
        } else {
            const float df = (a - b).f(), fb = b.f();  // a - b rounded to the half type first
            diff = (double)fabsf(df);
            abs_b = (double)fabsf(fb);
            both_nan = a.f() != a.f() && fb != fb;
        }
        const bool ok = diff <= d.fp1 + d.fp0 * abs_b || (d.ip0 && both_nan);
        return ok ? 1 : 0;
    }
    static __device__ __forceinline__ S comb(S a, S b) { return a & b; }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};

}  // namespace

#if RC_EXTX_KIND == 0
#define RC_EXTX_NAME(f) f##_h16
using ExtT = h16;
#elif RC_EXTX_KIND == 1
#define RC_EXTX_NAME(f) f##_b16
using ExtT = b16;
#elif RC_EXTX_KIND == 2
#define RC_EXTX_NAME(f) f##_c32
using ExtT = c32;
#else
#define RC_EXTX_NAME(f) f##_c64
using ExtT = c64;
#endif

void RC_EXTX_NAME(reduce_ext)(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    if constexpr (is_half_t<ExtT>::value) reduce_half<ExtT>(dev, op, cr, a, out, n);
    else reduce_cplx<typename real_of<ExtT>::type>(dev, op, cr, a, out, n);
}
void RC_EXTX_NAME(vecdot_ext)(rc_device *dev, const CanonRed &cr, const void *a, const void *b, void *c, int64_t n) {
    if constexpr (is_half_t<ExtT>::value) reduce_typed<PHDot<ExtT>>(dev, cr, a, c, n, b, 0.0, 0.0, 0);
    else reduce_typed<PCDot<typename real_of<ExtT>::type>>(dev, cr, a, c, n, b, 0.0, 0.0, 0);
}
void RC_EXTX_NAME(allclose_ext)(rc_device *dev, const CanonRed &cr, const void *a, const void *b, void *out, int64_t n,
                                double rtol, double atol, int equal_nan) {
    reduce_typed<PCloseX<ExtT>>(dev, cr, a, out, n, b, rtol, atol, equal_nan);
}

}  // namespace rc
