// rc_ew_ext.cuh -- functors of the extended element types (f16, bf16, c32, c64; rc_types.cuh), shared by the
// rc_ew_ext_*.cu translation units (split so that they build in parallel).
//   half     FViaF32<T, F>: both operands to f32, the f32 functor, ONE rounding back -- `half` crate semantics, and exactly
//            what NumPy's float16 / ml_dtypes' bfloat16 do, so + - * / are bit-exact against them;
//   complex  + - * / neg from the operators of rc_types.cuh (num-complex formulas); abs = hypot(re, im) (`norm()`),
//            real / imag with REAL output, conj, square = z * z, reciprocal = (re / n, -im / n) (`inv()`); the
//            transcendental functions through thrust::complex (tolerance-compared, like every libm function here).
#pragma once
#include <thrust/complex.h>

#include "rc_dispatch.cuh"
#include "rc_types.cuh"

namespace rc {
namespace {

// ---------------- half: everything through f32 ----------------
template <class T, template <class> class FF>
struct FViaF32 { using TA = T; using TB = T; using TO = T; static constexpr int NIN = FF<float>::NIN;
    RC_FN T apply(T a) { return T(FF<float>::apply(a.f())); }
    RC_FN T apply(T a, T b) { return T(FF<float>::apply(a.f(), b.f())); } };
template <class T, template <class> class FF>
struct FViaF32Bool { using TA = T; using TB = T; using TO = uint8_t; static constexpr int NIN = FF<float>::NIN;
    RC_FN uint8_t apply(T a) { return FF<float>::apply(a.f()); }
    RC_FN uint8_t apply(T a, T b) { return FF<float>::apply(a.f(), b.f()); } };

// ---------------- complex ----------------
template <class R> struct FCAbs { using TA = cplx<R>; using TB = TA; using TO = R; static constexpr int NIN = 1;
    RC_FN R apply(TA a) { if constexpr (sizeof(R) == 4) return hypotf(a.re, a.im); else return hypot(a.re, a.im); } };
template <class R> struct FCReal { using TA = cplx<R>; using TB = TA; using TO = R; static constexpr int NIN = 1; RC_FN R apply(TA a) { return a.re; } };
template <class R> struct FCImag { using TA = cplx<R>; using TB = TA; using TO = R; static constexpr int NIN = 1; RC_FN R apply(TA a) { return a.im; } };
template <class R> struct FCConj { using TA = cplx<R>; using TB = TA; using TO = TA; static constexpr int NIN = 1; RC_FN TA apply(TA a) { return TA(a.re, -a.im); } };
template <class R> struct FCRecip { using TA = cplx<R>; using TB = TA; using TO = TA; static constexpr int NIN = 1;
    RC_FN TA apply(TA a) { const R n = a.re * a.re + a.im * a.im; return TA(a.re / n, -a.im / n); } };
#define RC_CPLX_MATH(NAME, EXPR)                                                                             \
    template <class R> struct NAME { using TA = cplx<R>; using TB = TA; using TO = TA; static constexpr int NIN = 1; \
        RC_FN TA apply(TA a) { const thrust::complex<R> z(a.re, a.im); const thrust::complex<R> r = EXPR; return TA(r.real(), r.imag()); } };
RC_CPLX_MATH(FCExp, thrust::exp(z))
RC_CPLX_MATH(FCLog, thrust::log(z))
RC_CPLX_MATH(FCSqrt, thrust::sqrt(z))
RC_CPLX_MATH(FCSin, thrust::sin(z))
RC_CPLX_MATH(FCCos, thrust::cos(z))
RC_CPLX_MATH(FCSinh, thrust::sinh(z))
RC_CPLX_MATH(FCCosh, thrust::cosh(z))
RC_CPLX_MATH(FCTanh, thrust::tanh(z))
RC_CPLX_MATH(FCTan, thrust::tan(z))
RC_CPLX_MATH(FCAsin, thrust::asin(z))
RC_CPLX_MATH(FCAcos, thrust::acos(z))
RC_CPLX_MATH(FCAtan, thrust::atan(z))
RC_CPLX_MATH(FCAsinh, thrust::asinh(z))
RC_CPLX_MATH(FCAcosh, thrust::acosh(z))
RC_CPLX_MATH(FCAtanh, thrust::atanh(z))
// num-complex: log2 / log10 = ln(z) scaled by 1 / ln(base) on both components (Complex::log(base) via to_polar)
RC_CPLX_MATH(FCLog2, thrust::log(z) / thrust::complex<R>((R)0.693147180559945309417232121458176568))
RC_CPLX_MATH(FCLog10, thrust::log(z) / thrust::complex<R>((R)2.302585092994045684017991454684364208))
// ext_sign of a complex number: z / |z| (componentwise division by the real norm), zero for a zero magnitude
// (rstsr-dtype-traits/src/ext_num.rs:268-280)
template <class R> struct FCSign { using TA = cplx<R>; using TB = TA; using TO = TA; static constexpr int NIN = 1;
    RC_FN TA apply(TA a) {
        R n;
        if constexpr (sizeof(R) == 4) n = hypotf(a.re, a.im); else n = hypot(a.re, a.im);
        return n == (R)0 ? TA((R)0, (R)0) : TA(a.re / n, a.im / n); } };
// elementwise isclose for half / complex (isclose.rs:92-106, TE = f64): |a - b| and |b| in the element type, then widened
template <class T> struct FIsCloseX { using TA = T; using TB = T; using TO = uint8_t; static constexpr int NIN = 2;
    using Params = IsCloseParams;
    RC_FN uint8_t apply(T a, T b, const IsCloseParams &p) {
        double diff, abs_b;
        bool both_nan;
        if constexpr (is_cplx_t<T>::value) {
            using R = typename real_of<T>::type;
            const T df = a - b;
            if constexpr (sizeof(R) == 4) { diff = (double)hypotf(df.re, df.im); abs_b = (double)hypotf(b.re, b.im); }
            else { diff = hypot(df.re, df.im); abs_b = hypot(b.re, b.im); }
            both_nan = (a.re != a.re || a.im != a.im) && (b.re != b.re || b.im != b.im);
        } else {
            const float df = (a - b).f(), fb = b.f();
            diff = (double)fabsf(df);
            abs_b = (double)fabsf(fb);
            both_nan = a.f() != a.f() && fb != fb;
        }
        return (diff <= p.atol + p.rtol * abs_b || (p.equal_nan && both_nan)) ? 1 : 0; } };
// predicates (num-complex): is_nan = re or im NaN; is_infinite = not NaN and re or im infinite; is_finite = both finite
template <class R> struct FCIsNan { using TA = cplx<R>; using TB = TA; using TO = uint8_t; static constexpr int NIN = 1;
    RC_FN uint8_t apply(TA a) { return (a.re != a.re) || (a.im != a.im); } };
template <class R> struct FCIsInf { using TA = cplx<R>; using TB = TA; using TO = uint8_t; static constexpr int NIN = 1;
    RC_FN uint8_t apply(TA a) { return !((a.re != a.re) || (a.im != a.im)) && (isinf(a.re) || isinf(a.im)); } };
template <class R> struct FCIsFinite { using TA = cplx<R>; using TB = TA; using TO = uint8_t; static constexpr int NIN = 1;
    RC_FN uint8_t apply(TA a) { return isfinite(a.re) && isfinite(a.im); } };

// ---------------- casts ----------------
template <class TOut, class TIn> struct conv_t {
    RC_FN TOut go(TIn a) {
        if constexpr (is_half_t<TIn>::value) {                       // half -> f32 / f64 / bool / other half
            if constexpr (std::is_same<TOut, uint8_t>::value) return a.f() != 0.0f ? 1 : 0;
            else if constexpr (is_half_t<TOut>::value) return TOut(a.f());
            else return (TOut)a.f();
        } else if constexpr (is_half_t<TOut>::value) {               // f32 / f64 / bool -> half: one rounding
            return TOut(a);
        } else if constexpr (is_cplx_t<TIn>::value) {                // c32 <-> c64: componentwise `as`
            using RO = typename real_of<TOut>::type;
            return TOut((RO)a.re, (RO)a.im);
        } else {                                                     // real -> complex: (a as R, 0)
            using RO = typename real_of<TOut>::type;
            return TOut((RO)a, (RO)0);
        }
    }
};
template <class TOut, class TIn> struct FCastX { using TA = TIn; using TB = TIn; using TO = TOut; static constexpr int NIN = 1;
    RC_FN TOut apply(TIn a) { return conv_t<TOut, TIn>::go(a); } };
struct BoolIn { uint8_t v; };  // bool source: 0 / 1
template <class TOut> struct FCastFromBool { using TA = uint8_t; using TB = uint8_t; using TO = TOut; static constexpr int NIN = 1;
    RC_FN TOut apply(uint8_t a) { return TOut(a ? 1.0f : 0.0f); } };

struct alignas(16) U128 { uint64_t lo, hi; };

}  // namespace
}  // namespace rc
