// rc_types.cuh -- the element types beyond the Rust primitives: IEEE half, bfloat16, complex<f32>, complex<f64>
// (`half::f16`, `half::bf16`, `num::Complex<f32>`, `num::Complex<f64>`; umbrella bound
// DeviceComplexFloatAPI, rstsr-core/src/operators/combined_trait.rs:6-55).
//
// Semantics follow the crates the reference uses, so results can be compared bit for bit where IEEE allows:
//   half::f16 / bf16   every arithmetic op converts both operands to f32, computes there and rounds ONCE back
//                      (`impl Add for f16 { f16::from_f32(f32::from(self) + f32::from(rhs)) }`), math functions likewise
//                      (num_traits::Float for f16 goes through f32); conversions are single roundings to nearest even
//                      (from_f32 / from_f64).  NumPy's float16 and ml_dtypes' bfloat16 do the same.
//   num::Complex<T>    add / sub componentwise; mul = (ac - bd, ad + bc); div = ((ac + bd) / n, (bc - ad) / n) with
//                      n = c^2 + d^2 -- the textbook formulas, no scaling (num-complex `impl Div`), so -fmad=false code
//                      reproduces them exactly; norm() = hypot(re, im).
// All four are plain-old-data (2, 2, 8, 16 bytes) and move through the copy / gather kernels as raw words.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <limits>
#include <type_traits>

namespace rc {

#define RC_HD __host__ __device__ __forceinline__

struct h16 {
    uint16_t bits;
    h16() = default;
    RC_HD explicit h16(float f) { __half h = __float2half_rn(f); bits = *reinterpret_cast<uint16_t *>(&h); }
    RC_HD explicit h16(double d) { __half h = __double2half(d); bits = *reinterpret_cast<uint16_t *>(&h); }
    RC_HD explicit h16(int v) : h16((float)v) {}
    RC_HD explicit h16(unsigned int v) : h16((double)v) {}
    RC_HD explicit h16(long long v) : h16((double)v) {}
    RC_HD explicit h16(long v) : h16((double)v) {}
    RC_HD explicit h16(unsigned long v) : h16((double)v) {}
    RC_HD explicit h16(unsigned long long v) : h16((double)v) {}
    RC_HD float f() const { __half h; *reinterpret_cast<uint16_t *>(&h) = bits; return __half2float(h); }
    RC_HD explicit operator float() const { return f(); }
    RC_HD explicit operator double() const { return (double)f(); }
};

struct b16 {
    uint16_t bits;
    b16() = default;
    RC_HD explicit b16(float f) { __nv_bfloat16 h = __float2bfloat16_rn(f); bits = *reinterpret_cast<uint16_t *>(&h); }
    RC_HD explicit b16(double d) { __nv_bfloat16 h = __double2bfloat16(d); bits = *reinterpret_cast<uint16_t *>(&h); }
    RC_HD explicit b16(int v) : b16((float)v) {}
    RC_HD explicit b16(unsigned int v) : b16((double)v) {}
    RC_HD explicit b16(long long v) : b16((double)v) {}
    RC_HD explicit b16(long v) : b16((double)v) {}
    RC_HD explicit b16(unsigned long v) : b16((double)v) {}
    RC_HD explicit b16(unsigned long long v) : b16((double)v) {}
    RC_HD float f() const { return __uint_as_float_hd((uint32_t)bits << 16); }
    RC_HD explicit operator float() const { return f(); }
    RC_HD explicit operator double() const { return (double)f(); }

  private:
    static RC_HD float __uint_as_float_hd(uint32_t u) {
        float r;
#ifdef __CUDA_ARCH__
        r = __uint_as_float(u);
#else
        static_assert(sizeof(float) == 4, "float");
        __builtin_memcpy(&r, &u, 4);
#endif
        return r;
    }
};

template <class T> struct is_half_t : std::integral_constant<bool, std::is_same<T, h16>::value || std::is_same<T, b16>::value> {};

// every op: f32 compute, one rounding
#define RC_HALF_BINOP(OP)                                                                         \
    RC_HD h16 operator OP(h16 a, h16 b) { return h16(a.f() OP b.f()); }                           \
    RC_HD b16 operator OP(b16 a, b16 b) { return b16(a.f() OP b.f()); }
RC_HALF_BINOP(+)
RC_HALF_BINOP(-)
RC_HALF_BINOP(*)
RC_HALF_BINOP(/)
#undef RC_HALF_BINOP
RC_HD h16 operator-(h16 a) { h16 r; r.bits = a.bits ^ 0x8000u; return r; }
RC_HD b16 operator-(b16 a) { b16 r; r.bits = a.bits ^ 0x8000u; return r; }
#define RC_HALF_CMP(OP)                                                     \
    RC_HD bool operator OP(h16 a, h16 b) { return a.f() OP b.f(); }         \
    RC_HD bool operator OP(b16 a, b16 b) { return a.f() OP b.f(); }
RC_HALF_CMP(==)
RC_HALF_CMP(!=)
RC_HALF_CMP(<)
RC_HALF_CMP(<=)
RC_HALF_CMP(>)
RC_HALF_CMP(>=)
#undef RC_HALF_CMP

template <class R>
struct cplx {
    R re, im;
    cplx() = default;
    RC_HD cplx(R r, R i) : re(r), im(i) {}
    RC_HD explicit cplx(R r) : re(r), im((R)0) {}
    RC_HD explicit cplx(int v) : re((R)v), im((R)0) {}
    RC_HD explicit cplx(unsigned int v) : re((R)v), im((R)0) {}
    RC_HD explicit cplx(long long v) : re((R)v), im((R)0) {}
    RC_HD explicit cplx(long v) : re((R)v), im((R)0) {}
    RC_HD explicit cplx(unsigned long v) : re((R)v), im((R)0) {}
    RC_HD explicit cplx(unsigned long long v) : re((R)v), im((R)0) {}
};
using c32 = cplx<float>;
using c64 = cplx<double>;
static_assert(sizeof(h16) == 2 && sizeof(b16) == 2 && sizeof(c32) == 8 && sizeof(c64) == 16, "element sizes");

template <class T> struct is_cplx_t : std::false_type {};
template <class R> struct is_cplx_t<cplx<R>> : std::true_type {};
template <class T> struct real_of { using type = T; };
template <class R> struct real_of<cplx<R>> { using type = R; };

template <class R> RC_HD cplx<R> operator+(cplx<R> a, cplx<R> b) { return cplx<R>(a.re + b.re, a.im + b.im); }
template <class R> RC_HD cplx<R> operator-(cplx<R> a, cplx<R> b) { return cplx<R>(a.re - b.re, a.im - b.im); }
template <class R> RC_HD cplx<R> operator-(cplx<R> a) { return cplx<R>(-a.re, -a.im); }
template <class R> RC_HD cplx<R> operator*(cplx<R> a, cplx<R> b) {
    return cplx<R>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <class R> RC_HD cplx<R> operator/(cplx<R> a, cplx<R> b) {
    const R n = b.re * b.re + b.im * b.im;
    return cplx<R>((a.re * b.re + a.im * b.im) / n, (a.im * b.re - a.re * b.im) / n);
}
template <class R> RC_HD cplx<R> operator/(cplx<R> a, R s) { return cplx<R>(a.re / s, a.im / s); }
template <class R> RC_HD bool operator==(cplx<R> a, cplx<R> b) { return a.re == b.re && a.im == b.im; }
template <class R> RC_HD bool operator!=(cplx<R> a, cplx<R> b) { return !(a == b); }

}  // namespace rc

// finite extremes for max / min reductions over half types (ExtReal::ext_min_value / ext_max_value = MIN / MAX)
namespace std {
template <> struct numeric_limits<rc::h16> {
    static RC_HD rc::h16 lowest() { rc::h16 r; r.bits = 0xFBFFu; return r; }
    static RC_HD rc::h16 max() { rc::h16 r; r.bits = 0x7BFFu; return r; }
};
template <> struct numeric_limits<rc::b16> {
    static RC_HD rc::b16 lowest() { rc::b16 r; r.bits = 0xFF7Fu; return r; }
    static RC_HD rc::b16 max() { rc::b16 r; r.bits = 0x7F7Fu; return r; }
};
}  // namespace std
