// rc_ew_ext_cplx_math2.cu -- the inverse trigonometric / hyperbolic functions of c32 / c64 through thrust::complex (see rc_ew_ext.cuh).
// Flat kernels only (ALLOW_TILE = false): these functors are 100-400 instructions each and dominate the library's compile
// time; a transposed operand of a transcendental function is rare and still correct through the flat kernel.
#include "rc_ew_ext.cuh"

namespace rc {

bool run_unary_cplx_math2(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
#define RC_CPLX_UN(OPCODE, FF)                                                         \
    case OPCODE:                                                                       \
        if (t == RC_C32) { ew_launch<FF<float>, false>(dev, c, args); return true; }    \
        if (t == RC_C64) { ew_launch<FF<double>, false>(dev, c, args); return true; }   \
        break;
    switch (op) {
        RC_CPLX_UN(RC_ASIN, FCAsin) RC_CPLX_UN(RC_ACOS, FCAcos) RC_CPLX_UN(RC_ATAN, FCAtan)
        RC_CPLX_UN(RC_ASINH, FCAsinh) RC_CPLX_UN(RC_ACOSH, FCAcosh) RC_CPLX_UN(RC_ATANH, FCAtanh)
        default: break;
    }
#undef RC_CPLX_UN
    return false;
}

}  // namespace rc
