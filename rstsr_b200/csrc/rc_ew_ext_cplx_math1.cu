// rc_ew_ext_cplx_math1.cu -- exp / log / sqrt and the direct trigonometric / hyperbolic functions of c32 / c64 through thrust::complex (see rc_ew_ext.cuh).
// Flat kernels only (ALLOW_TILE = false): these functors are 100-400 instructions each and dominate the library's compile
// time; a transposed operand of a transcendental function is rare and still correct through the flat kernel.
#include "rc_ew_ext.cuh"

namespace rc {

bool run_unary_cplx_math1(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
#define RC_CPLX_UN(OPCODE, FF)                                                         \
    case OPCODE:                                                                       \
        if (t == RC_C32) { ew_launch<FF<float>, false>(dev, c, args); return true; }    \
        if (t == RC_C64) { ew_launch<FF<double>, false>(dev, c, args); return true; }   \
        break;
    switch (op) {
        RC_CPLX_UN(RC_EXP, FCExp) RC_CPLX_UN(RC_LOG, FCLog) RC_CPLX_UN(RC_LOG2, FCLog2) RC_CPLX_UN(RC_LOG10, FCLog10)
        RC_CPLX_UN(RC_SQRT, FCSqrt) RC_CPLX_UN(RC_SIN, FCSin) RC_CPLX_UN(RC_COS, FCCos) RC_CPLX_UN(RC_TAN, FCTan)
        RC_CPLX_UN(RC_SINH, FCSinh) RC_CPLX_UN(RC_COSH, FCCosh) RC_CPLX_UN(RC_TANH, FCTanh)
        default: break;
    }
#undef RC_CPLX_UN
    return false;
}

}  // namespace rc
