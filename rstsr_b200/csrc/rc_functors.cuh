// rc_functors.cuh -- per-element semantics of every elementwise op, matching the reference closures:
//   arithmetic  rstsr-core/src/feature_rayon/auto_impl/op_ternary_arithmetic.rs:3-15  (c = a o b)
//   functions   auto_impl/op_ternary_common.rs (maximum/minimum/floor_divide/pow/atan2/..., comparisons)
//   unary       auto_impl/op_binary_arithmetic.rs:94-113 (neg/not), auto_impl/op_binary_common.rs
//   casts       rstsr-dtype-traits/src/promotion.rs (Rust `as`; to bool is `!= 0`)
//   min/max     rstsr-dtype-traits/src/ext_real.rs:32-87 (floats: NaN-ignoring f64::max/min)
// Integer arithmetic wraps (the reference's CI runs --release).  Where Rust would panic (integer
// division by zero, MIN / -1) a GPU kernel cannot: the result is 0 (x / 0, x % 0) resp. the wrapped value.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

namespace rc {

struct bool_t {  // Rust bool: one byte holding 0 or 1
    uint8_t v;
};

template <class T> struct is_float_t : std::integral_constant<bool, std::is_floating_point<T>::value> {};
template <class T> using uns_t = typename std::make_unsigned<T>::type;

#define RC_FN static __device__ __forceinline__

// ---------------- binary: TA = TB = TO = T ----------------
template <class T> struct FAdd { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) { if constexpr (std::is_integral<T>::value) return (T)((uns_t<T>)a + (uns_t<T>)b); else return a + b; } };
template <class T> struct FSub { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) { if constexpr (std::is_integral<T>::value) return (T)((uns_t<T>)a - (uns_t<T>)b); else return a - b; } };
template <class T> struct FMul { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) {
        if constexpr (std::is_integral<T>::value) {
            if constexpr (sizeof(T) < 4) return (T)((unsigned)a * (unsigned)b); else return (T)((uns_t<T>)a * (uns_t<T>)b);
        } else return a * b; } };
template <class T> struct FDiv { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) {
        if constexpr (std::is_integral<T>::value) {
            if (b == 0) return 0;
            if constexpr (std::is_signed<T>::value) { if (b == (T)-1) return (T)(0 - (uns_t<T>)a); }
            return a / b;
        } else return a / b; } };
template <class T> struct FRem { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) {
        if constexpr (std::is_integral<T>::value) {
            if (b == 0) return 0;
            if constexpr (std::is_signed<T>::value) { if (b == (T)-1) return 0; }
            return a % b;
        } else if constexpr (sizeof(T) == 4) return fmodf(a, b); else return fmod(a, b); } };
template <class T> struct FBitOr { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2; RC_FN T apply(T a, T b) { return (T)(a | b); } };
template <class T> struct FBitAnd { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2; RC_FN T apply(T a, T b) { return (T)(a & b); } };
template <class T> struct FBitXor { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2; RC_FN T apply(T a, T b) { return (T)(a ^ b); } };
// release-mode Rust masks the shift amount to the bit width
template <class T> struct FShl { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) { return (T)((uns_t<T>)a << ((unsigned)b & (8 * sizeof(T) - 1))); } };
template <class T> struct FShr { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) { return (T)(a >> ((unsigned)b & (8 * sizeof(T) - 1))); } };
template <class T> struct FMaximum { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) {
        if constexpr (std::is_same<T, float>::value) return fmaxf(a, b);
        else if constexpr (std::is_same<T, double>::value) return fmax(a, b);
        else return a < b ? b : a; } };
template <class T> struct FMinimum { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) {
        if constexpr (std::is_same<T, float>::value) return fminf(a, b);
        else if constexpr (std::is_same<T, double>::value) return fmin(a, b);
        else return b < a ? b : a; } };
template <class T> struct FFloorDivide { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, T b) {
        if constexpr (std::is_same<T, float>::value) return floorf(a / b);
        else if constexpr (std::is_same<T, double>::value) return floor(a / b);
        else {
            if (b == 0) return 0;
            if constexpr (std::is_signed<T>::value) {
                if (b == (T)-1) return (T)(0 - (uns_t<T>)a);
                T q = a / b, r = a % b;
                return (r != 0 && ((r < 0) != (b < 0))) ? (T)(q - 1) : q;
            } else return a / b;
        } } };
#define RC_FLOAT_BINARY(NAME, F32, F64)                                                        \
    template <class T> struct NAME { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 2; \
        RC_FN T apply(T a, T b) { if constexpr (sizeof(T) == 4) return F32; else return F64; } };
RC_FLOAT_BINARY(FPow, powf(a, b), pow(a, b))
RC_FLOAT_BINARY(FAtan2, atan2f(a, b), atan2(a, b))
RC_FLOAT_BINARY(FCopysign, copysignf(a, b), copysign(a, b))
RC_FLOAT_BINARY(FHypot, hypotf(a, b), hypot(a, b))
RC_FLOAT_BINARY(FLogAddExp, logf(expf(a) + expf(b)), log(exp(a) + exp(b)))
RC_FLOAT_BINARY(FNextAfter, nextafterf(a, b), nextafter(a, b))

// pow with a mixed exponent (num::Pow<TB> for TA, the bound of OpPowAPI: auto_impl/op_ternary_common.rs:141-149):
//   float ^ i8/u8/i16/u16/i32 = powi: compiler-rt's __powi?f2 (square-and-multiply from the low bit, reciprocal at the
//   end for negative exponents) -- restated so the result is bit-identical to Rust's f32::powi / f64::powi;
//   int ^ u8/u16/u32/u64 = wrapping power (release-mode {integer}::pow); the exponent arrives as u32.
template <class T> struct FPowi { using TA = T; using TB = int32_t; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, int32_t b) {
        const bool recip = b < 0;
        T r = (T)1;
        while (true) {
            if (b & 1) r *= a;
            b /= 2;
            if (b == 0) break;
            a *= a;
        }
        return recip ? (T)1 / r : r; } };
template <class T> struct FIPow { using TA = T; using TB = uint32_t; using TO = T; static constexpr int NIN = 2;
    RC_FN T apply(T a, uint32_t e) {
        using U = typename std::conditional<(sizeof(T) < 4), uint32_t, uns_t<T>>::type;
        U base = (U)(uns_t<T>)a, r = 1;
        while (e) {
            if (e & 1u) r *= base;
            base *= base;
            e >>= 1;
        }
        return (T)(uns_t<T>)r; } };

// isclose (rstsr-dtype-traits/src/isclose.rs:92-106) with TE = f64: diff and |b| are formed in the element type,
// then cast; inf vs inf gives |inf - inf| = NaN -> not close, as in the reference.
struct IsCloseParams { double rtol, atol; int equal_nan; };
template <class T> struct FIsClose { using TA = T; using TB = T; using TO = uint8_t; static constexpr int NIN = 2;
    using Params = IsCloseParams;
    RC_FN uint8_t apply(T a, T b, const IsCloseParams &p) {
        double diff, abs_b;
        if constexpr (std::is_floating_point<T>::value) {
            const T df = a - b;
            diff = (double)(df < (T)0 ? -df : df);
            abs_b = (double)(b < (T)0 ? -b : b);
            if (df != df) diff = (double)df;
            if (b != b) abs_b = (double)b;
        } else if constexpr (std::is_signed<T>::value) {
            const T df = a >= b ? (T)((uns_t<T>)a - (uns_t<T>)b) : (T)((uns_t<T>)b - (uns_t<T>)a);
            diff = (double)df;
            abs_b = (double)(b < 0 ? (T)((uns_t<T>)0 - (uns_t<T>)b) : b);
        } else {
            diff = (double)(a >= b ? (T)(a - b) : (T)(b - a));
            abs_b = (double)b;
        }
        bool ok = diff <= p.atol + p.rtol * abs_b;
        if constexpr (std::is_floating_point<T>::value) ok = ok || (p.equal_nan && a != a && b != b);
        return ok ? 1 : 0; } };

// ---------------- comparisons: TO = bool ----------------
#define RC_COMPARE(NAME, EXPR)                                                                  \
    template <class T> struct NAME { using TA = T; using TB = T; using TO = uint8_t; static constexpr int NIN = 2; \
        RC_FN uint8_t apply(T a, T b) { return (EXPR) ? 1 : 0; } };
RC_COMPARE(FEq, a == b)
RC_COMPARE(FNe, a != b)
RC_COMPARE(FLt, a < b)
RC_COMPARE(FLe, a <= b)
RC_COMPARE(FGt, a > b)
RC_COMPARE(FGe, a >= b)

// ---------------- unary ----------------
struct NoOperand {};
#define RC_UNARY_HEAD(TOUT) using TA = T; using TB = T; using TO = TOUT; static constexpr int NIN = 1;
template <class T> struct FNeg { RC_UNARY_HEAD(T)
    RC_FN T apply(T a) { if constexpr (std::is_integral<T>::value) return (T)(0 - (uns_t<T>)a); else return -a; } };
template <class T> struct FNot { RC_UNARY_HEAD(T) RC_FN T apply(T a) { return (T)~a; } };
struct FNotBool { using T = uint8_t; RC_UNARY_HEAD(uint8_t) RC_FN uint8_t apply(uint8_t a) { return a ? 0 : 1; } };
template <class T> struct FAbs { RC_UNARY_HEAD(T)
    RC_FN T apply(T a) {
        if constexpr (std::is_same<T, float>::value) return fabsf(a);
        else if constexpr (std::is_same<T, double>::value) return fabs(a);
        else if constexpr (std::is_signed<T>::value) return a < 0 ? (T)(0 - (uns_t<T>)a) : a;
        else return a; } };
template <class T> struct FSquare { RC_UNARY_HEAD(T) RC_FN T apply(T a) { return FMul<T>::apply(a, a); } };
template <class T> struct FSign { RC_UNARY_HEAD(T)
    RC_FN T apply(T a) {
        if constexpr (std::is_floating_point<T>::value) { if (a != a) return a; return a > 0 ? (T)1 : (a < 0 ? (T)-1 : (T)0); }
        else if constexpr (std::is_signed<T>::value) return a > 0 ? (T)1 : (a < 0 ? (T)-1 : (T)0);
        else return a == 0 ? (T)0 : (T)1; } };
#define RC_FLOAT_UNARY(NAME, F32, F64)                                                         \
    template <class T> struct NAME { RC_UNARY_HEAD(T)                                          \
        RC_FN T apply(T a) { if constexpr (sizeof(T) == 4) return F32; else return F64; } };
RC_FLOAT_UNARY(FSqrt, sqrtf(a), sqrt(a))
RC_FLOAT_UNARY(FExp, expf(a), exp(a))
RC_FLOAT_UNARY(FExpm1, expm1f(a), expm1(a))
RC_FLOAT_UNARY(FLog, logf(a), log(a))
RC_FLOAT_UNARY(FLog2, log2f(a), log2(a))
RC_FLOAT_UNARY(FLog10, log10f(a), log10(a))
RC_FLOAT_UNARY(FSin, sinf(a), sin(a))
RC_FLOAT_UNARY(FCos, cosf(a), cos(a))
RC_FLOAT_UNARY(FTan, tanf(a), tan(a))
RC_FLOAT_UNARY(FAsin, asinf(a), asin(a))
RC_FLOAT_UNARY(FAcos, acosf(a), acos(a))
RC_FLOAT_UNARY(FAtan, atanf(a), atan(a))
RC_FLOAT_UNARY(FSinh, sinhf(a), sinh(a))
RC_FLOAT_UNARY(FCosh, coshf(a), cosh(a))
RC_FLOAT_UNARY(FTanh, tanhf(a), tanh(a))
RC_FLOAT_UNARY(FAsinh, asinhf(a), asinh(a))
RC_FLOAT_UNARY(FAcosh, acoshf(a), acosh(a))
RC_FLOAT_UNARY(FAtanh, atanhf(a), atanh(a))
RC_FLOAT_UNARY(FFloor, floorf(a), floor(a))
RC_FLOAT_UNARY(FCeil, ceilf(a), ceil(a))
RC_FLOAT_UNARY(FRound, roundf(a), round(a))
RC_FLOAT_UNARY(FTrunc, truncf(a), trunc(a))
RC_FLOAT_UNARY(FRecip, 1.0f / a, 1.0 / a)
template <class T> struct FIdentity { RC_UNARY_HEAD(T) RC_FN T apply(T a) { return a; } };
template <class T> struct FZero { RC_UNARY_HEAD(T) RC_FN T apply(T) { return (T)0; } };
#define RC_PREDICATE(NAME, EXPR)                                                               \
    template <class T> struct NAME { RC_UNARY_HEAD(uint8_t) RC_FN uint8_t apply(T a) { return (EXPR) ? 1 : 0; } };
RC_PREDICATE(FIsNan, a != a)
RC_PREDICATE(FIsInf, isinf(a))
RC_PREDICATE(FIsFinite, isfinite(a))
// OpSignBitAPI in the reference writes `b.is_positive()` (auto_impl/op_binary_common.rs:104) -- for floats
// that is "sign bit clear".  Kept as is: a drop-in device returns what the reference device returns.
RC_PREDICATE(FSignBit, !signbit(a))

// ---------------- casts (assign with promotion, fill) ----------------
template <class TOut, class TIn, bool OUT_BOOL, bool IN_BOOL>
struct FCast { using TA = TIn; using TB = TIn; using TO = TOut; static constexpr int NIN = 1;
    RC_FN TOut apply(TIn a) {
        if constexpr (OUT_BOOL) return (TOut)(a != (TIn)0 ? 1 : 0);   // `self != 0`
        else if constexpr (IN_BOOL) return (TOut)(a ? 1 : 0);
        else if constexpr (std::is_floating_point<TIn>::value && std::is_integral<TOut>::value) {
            // Rust `as`: saturating, NaN -> 0.  cvt.rzi saturates for 32/64-bit targets; narrow targets clamp here.
            if (a != a) return (TOut)0;
            if constexpr (sizeof(TOut) < 4) {
                const TIn lo = (TIn)std::numeric_limits<TOut>::min(), hi = (TIn)std::numeric_limits<TOut>::max();
                TIn t = a < lo ? lo : (a > hi ? hi : a);
                return (TOut)(int)t;
            } else if constexpr (std::is_same<TOut, int32_t>::value) {
                if constexpr (sizeof(TIn) == 4) return __float2int_rz(a); else return __double2int_rz(a);
            } else if constexpr (std::is_same<TOut, uint32_t>::value) {
                if constexpr (sizeof(TIn) == 4) return __float2uint_rz(a); else return __double2uint_rz(a);
            } else if constexpr (std::is_same<TOut, int64_t>::value) {
                if constexpr (sizeof(TIn) == 4) return __float2ll_rz(a); else return __double2ll_rz(a);
            } else {
                if constexpr (sizeof(TIn) == 4) return __float2ull_rz(a); else return __double2ull_rz(a);
            }
        } else return (TOut)a; } };

// fill: c = constant (NIN = 0; the constant arrives in the `a` scalar slot already cast on the host)
template <class T> struct FFill { using TA = T; using TB = T; using TO = T; static constexpr int NIN = 0;
    RC_FN T apply(T a) { return a; } };

}  // namespace rc
