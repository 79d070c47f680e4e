// rc_ew_ext.cu -- dispatch, casts, raw 16-byte moves and host scalar conversions of the extended element types
// (f16, bf16, c32, c64).  The op kernels live in rc_ew_ext_{half_bin,half_un,cplx}.cu; functors in rc_ew_ext.cuh.
#include "rc_ew_ext.cuh"

namespace rc {

bool run_binary_half(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
bool run_unary_half(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
bool run_binary_cplx(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
bool run_unary_cplx(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args);

// c = a o b / a = f(b) for the extended types (t = operand type); false when the op does not exist for t
bool run_binary_ext(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    if (dtype_is_half(t)) return run_binary_half(dev, op, t, c, args);
    if (dtype_is_complex(t)) return run_binary_cplx(dev, op, t, c, args);
    return false;
}
bool run_unary_ext(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    if (dtype_is_half(t)) return run_unary_half(dev, op, t, c, args);
    if (dtype_is_complex(t)) return run_unary_cplx(dev, op, t, c, args);
    return false;
}

// casts that involve an extended type; false = no such cast
bool run_cast_ext(rc_device *dev, rc_dtype tc, rc_dtype ta, const CanonEw &c, const EwArgs &args) {
#define RC_CX(TCODE, TOUT, ACODE, TIN) \
    if (tc == TCODE && ta == ACODE) { ew_launch<FCastX<TOUT, TIN>, false, false>(dev, c, args); return true; }
    // half <-> f32 / f64 / the other half
    RC_CX(RC_F16, h16, RC_F32, float) RC_CX(RC_F16, h16, RC_F64, double) RC_CX(RC_F16, h16, RC_BF16, b16)
    RC_CX(RC_BF16, b16, RC_F32, float) RC_CX(RC_BF16, b16, RC_F64, double) RC_CX(RC_BF16, b16, RC_F16, h16)
    RC_CX(RC_F32, float, RC_F16, h16) RC_CX(RC_F64, double, RC_F16, h16)
    RC_CX(RC_F32, float, RC_BF16, b16) RC_CX(RC_F64, double, RC_BF16, b16)
    RC_CX(RC_BOOL, uint8_t, RC_F16, h16) RC_CX(RC_BOOL, uint8_t, RC_BF16, b16)
    if (ta == RC_BOOL && tc == RC_F16) { ew_launch<FCastFromBool<h16>, false, false>(dev, c, args); return true; }
    if (ta == RC_BOOL && tc == RC_BF16) { ew_launch<FCastFromBool<b16>, false, false>(dev, c, args); return true; }
    // real -> complex, complex <-> complex
    RC_CX(RC_C32, c32, RC_F32, float) RC_CX(RC_C32, c32, RC_F64, double) RC_CX(RC_C32, c32, RC_C64, c64)
    RC_CX(RC_C64, c64, RC_F32, float) RC_CX(RC_C64, c64, RC_F64, double) RC_CX(RC_C64, c64, RC_C32, c32)
    RC_CX(RC_C64, c64, RC_I32, int32_t) RC_CX(RC_C64, c64, RC_I64, int64_t)
#undef RC_CX
    return false;
}

// 16-byte elements (c64) as raw words: copy and fill
void run_copy16(rc_device *dev, const CanonEw &c, const EwArgs &args) { ew_launch<FIdentity<U128>>(dev, c, args); }
void run_fill16(rc_device *dev, const CanonEw &c, const EwArgs &args) { ew_launch<FFill<U128>, false>(dev, c, args); }

// host scalar -> extended dtype (fill values, `numb` operands): src is f64 (re) or (re, im)
void host_to_ext(rc_dtype tc, double re, double im, void *out16) {
    std::memset(out16, 0, 16);
    switch (tc) {
        case RC_F16: { h16 v(re); std::memcpy(out16, &v, 2); return; }
        case RC_BF16: { b16 v(re); std::memcpy(out16, &v, 2); return; }
        case RC_C32: { c32 v((float)re, (float)im); std::memcpy(out16, &v, 8); return; }
        case RC_C64: { c64 v(re, im); std::memcpy(out16, &v, 16); return; }
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "not an extended dtype");
}
void host_from_ext(rc_dtype t, const void *src, double *re, double *im) {
    *im = 0.0;
    switch (t) {
        case RC_F16: { h16 v; std::memcpy(&v, src, 2); *re = (double)v.f(); return; }
        case RC_BF16: { b16 v; std::memcpy(&v, src, 2); *re = (double)v.f(); return; }
        case RC_C32: { c32 v; std::memcpy(&v, src, 8); *re = v.re; *im = v.im; return; }
        case RC_C64: { c64 v; std::memcpy(&v, src, 16); *re = v.re; *im = v.im; return; }
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "not an extended dtype");
}

}  // namespace rc
