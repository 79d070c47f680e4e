// rc_ew_ext_half_un.cu -- a = f(b) for f16 / bf16 (see rc_ew_ext.cuh)
#include "rc_ew_ext.cuh"

namespace rc {

bool run_unary_half(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
#define RC_HALF_UN(OPCODE, W, FF)                                              \
    case OPCODE:                                                               \
        if (t == RC_F16) { ew_launch<W<h16, FF>>(dev, c, args); return true; }  \
        if (t == RC_BF16) { ew_launch<W<b16, FF>>(dev, c, args); return true; } \
        break;
    if (dtype_is_half(t)) {
        switch (op) {
            RC_HALF_UN(RC_NEG, FViaF32, FNeg) RC_HALF_UN(RC_ABS, FViaF32, FAbs) RC_HALF_UN(RC_SQUARE, FViaF32, FSquare)
            RC_HALF_UN(RC_SIGN, FViaF32, FSign) RC_HALF_UN(RC_SQRT, FViaF32, FSqrt) RC_HALF_UN(RC_EXP, FViaF32, FExp)
            RC_HALF_UN(RC_EXPM1, FViaF32, FExpm1) RC_HALF_UN(RC_LOG, FViaF32, FLog) RC_HALF_UN(RC_LOG2, FViaF32, FLog2)
            RC_HALF_UN(RC_LOG10, FViaF32, FLog10) RC_HALF_UN(RC_SIN, FViaF32, FSin) RC_HALF_UN(RC_COS, FViaF32, FCos)
            RC_HALF_UN(RC_TAN, FViaF32, FTan) RC_HALF_UN(RC_ASIN, FViaF32, FAsin) RC_HALF_UN(RC_ACOS, FViaF32, FAcos)
            RC_HALF_UN(RC_ATAN, FViaF32, FAtan) RC_HALF_UN(RC_SINH, FViaF32, FSinh) RC_HALF_UN(RC_COSH, FViaF32, FCosh)
            RC_HALF_UN(RC_TANH, FViaF32, FTanh) RC_HALF_UN(RC_ASINH, FViaF32, FAsinh) RC_HALF_UN(RC_ACOSH, FViaF32, FAcosh)
            RC_HALF_UN(RC_ATANH, FViaF32, FAtanh) RC_HALF_UN(RC_FLOOR, FViaF32, FFloor) RC_HALF_UN(RC_CEIL, FViaF32, FCeil)
            RC_HALF_UN(RC_ROUND, FViaF32, FRound) RC_HALF_UN(RC_TRUNC, FViaF32, FTrunc) RC_HALF_UN(RC_RECIPROCAL, FViaF32, FRecip)
            RC_HALF_UN(RC_CONJ, FViaF32, FIdentity) RC_HALF_UN(RC_REAL, FViaF32, FIdentity) RC_HALF_UN(RC_IMAG, FViaF32, FZero)
            RC_HALF_UN(RC_ISNAN, FViaF32Bool, FIsNan) RC_HALF_UN(RC_ISINF, FViaF32Bool, FIsInf)
            RC_HALF_UN(RC_ISFINITE, FViaF32Bool, FIsFinite) RC_HALF_UN(RC_SIGNBIT, FViaF32Bool, FSignBit)
            default: break;
        }
        return false;
    }
#undef RC_HALF_UN
    return false;
}

}  // namespace rc
