// rc_reduce_extx.cu -- reductions of the extended element types (f16, bf16, c32, c64) on the kernels of rc_reduce.cuh.
//   half     every policy of the f32 path through PViaF32: inputs converted to f32, f32 state, ONE rounding of the result
//            (sum / prod / max / min / mean / var / std / l2_norm), index / count outputs unchanged (argmin / argmax /
//            count_nonzero).  The reference accumulates in the element type (`acc + x` in half precision): ours is the
//            more accurate value and lies within half-precision rounding of it.
//   complex  sum / prod / mean componentwise resp. with the operators of rc_types.cuh (the mean divides by
//            Complex::from(n)); var / std / l2_norm with REAL output as auto_impl/reduction.rs:207-354 defines them:
//            state (sum x, sum (x * conj x).re), var = q / n - (m * conj m).re with m = s / Complex::from(n).
#include "rc_canon.hpp"
#include "rc_device.hpp"

namespace rc {

#define RC_EXTX_DECL(sfx)                                                                                                      \
    void reduce_ext_##sfx(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n);                 \
    void vecdot_ext_##sfx(rc_device *dev, const CanonRed &cr, const void *a, const void *b, void *c, int64_t n);                  \
    void allclose_ext_##sfx(rc_device *dev, const CanonRed &cr, const void *a, const void *b, void *out, int64_t n, double rtol, \
                            double atol, int equal_nan);
RC_EXTX_DECL(h16) RC_EXTX_DECL(b16) RC_EXTX_DECL(c32) RC_EXTX_DECL(c64)
#undef RC_EXTX_DECL

void run_reduce_extx(rc_device *dev, rc_redop op, rc_dtype t, const CanonRed &cr, const void *a, void *out, int64_t n) {
    switch (t) {
        case RC_F16: reduce_ext_h16(dev, op, cr, a, out, n); return;
        case RC_BF16: reduce_ext_b16(dev, op, cr, a, out, n); return;
        case RC_C32: reduce_ext_c32(dev, op, cr, a, out, n); return;
        case RC_C64: reduce_ext_c64(dev, op, cr, a, out, n); return;
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "not an extended dtype");
}

void run_vecdot_extx(rc_device *dev, rc_dtype t, const CanonRed &cr, const void *a, const void *b, void *c, int64_t n) {
    switch (t) {
        case RC_F16: vecdot_ext_h16(dev, cr, a, b, c, n); return;
        case RC_BF16: vecdot_ext_b16(dev, cr, a, b, c, n); return;
        case RC_C32: vecdot_ext_c32(dev, cr, a, b, c, n); return;
        case RC_C64: vecdot_ext_c64(dev, cr, a, b, c, n); return;
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "not an extended dtype");
}

void run_allclose_extx(rc_device *dev, rc_dtype t, const CanonRed &cr, const void *a, const void *b, void *out, int64_t n,
                       double rtol, double atol, int equal_nan) {
    switch (t) {
        case RC_F16: allclose_ext_h16(dev, cr, a, b, out, n, rtol, atol, equal_nan); return;
        case RC_BF16: allclose_ext_b16(dev, cr, a, b, out, n, rtol, atol, equal_nan); return;
        case RC_C32: allclose_ext_c32(dev, cr, a, b, out, n, rtol, atol, equal_nan); return;
        case RC_C64: allclose_ext_c64(dev, cr, a, b, out, n, rtol, atol, equal_nan); return;
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "not an extended dtype");
}

}  // namespace rc
