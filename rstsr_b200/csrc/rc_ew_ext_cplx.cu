// rc_ew_ext_cplx.cu -- elementwise ops of c32 / c64 (see rc_ew_ext.cuh)
#include "rc_ew_ext.cuh"

namespace rc {

bool run_binary_cplx(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
#define RC_CPLX_OP(OPCODE, FF)                                                 \
    case OPCODE:                                                               \
        if (t == RC_C32) { ew_launch<FF<c32>>(dev, c, args); return true; }     \
        if (t == RC_C64) { ew_launch<FF<c64>>(dev, c, args); return true; }     \
        break;
    if (dtype_is_complex(t)) {
        switch (op) {
            RC_CPLX_OP(RC_ADD, FAdd) RC_CPLX_OP(RC_SUB, FSub) RC_CPLX_OP(RC_MUL, FMul) RC_CPLX_OP(RC_DIV, FDiv)
            RC_CPLX_OP(RC_EQ, FEq) RC_CPLX_OP(RC_NE, FNe)
            default: break;
        }
    }
#undef RC_CPLX_OP
    return false;
}

// thrust-based functions live in two translation units of their own (rc_ew_ext_cplx_math{1,2}.cu: they dominate the
// compile time) and skip the tile kernels: a transposed complex operand of a transcendental function takes the flat kernel
bool run_unary_cplx_math1(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
bool run_unary_cplx_math2(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args);

bool run_unary_cplx(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
#define RC_CPLX_UN(OPCODE, FF)                                                  \
    case OPCODE:                                                                \
        if (t == RC_C32) { ew_launch<FF<float>>(dev, c, args); return true; }    \
        if (t == RC_C64) { ew_launch<FF<double>>(dev, c, args); return true; }   \
        break;
#define RC_CPLX_UN_T(OPCODE, FF)                                                \
    case OPCODE:                                                                \
        if (t == RC_C32) { ew_launch<FF<c32>>(dev, c, args); return true; }      \
        if (t == RC_C64) { ew_launch<FF<c64>>(dev, c, args); return true; }      \
        break;
    if (dtype_is_complex(t)) {
        switch (op) {
            RC_CPLX_UN_T(RC_NEG, FNeg) RC_CPLX_UN_T(RC_SQUARE, FSquare)
            RC_CPLX_UN(RC_ABS, FCAbs) RC_CPLX_UN(RC_REAL, FCReal) RC_CPLX_UN(RC_IMAG, FCImag) RC_CPLX_UN(RC_CONJ, FCConj)
            RC_CPLX_UN(RC_RECIPROCAL, FCRecip) RC_CPLX_UN(RC_SIGN, FCSign)
            RC_CPLX_UN(RC_ISNAN, FCIsNan) RC_CPLX_UN(RC_ISINF, FCIsInf) RC_CPLX_UN(RC_ISFINITE, FCIsFinite)
            default: break;
        }
        return run_unary_cplx_math1(dev, op, t, c, args) || run_unary_cplx_math2(dev, op, t, c, args);
    }
#undef RC_CPLX_UN
#undef RC_CPLX_UN_T
    return false;
}

bool run_isclose_ext(rc_device *dev, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    switch (t) {
        case RC_F16: ew_launch<FIsCloseX<h16>>(dev, c, args); return true;
        case RC_BF16: ew_launch<FIsCloseX<b16>>(dev, c, args); return true;
        case RC_C32: ew_launch<FIsCloseX<c32>>(dev, c, args); return true;
        case RC_C64: ew_launch<FIsCloseX<c64>>(dev, c, args); return true;
        default: return false;
    }
}

}  // namespace rc
