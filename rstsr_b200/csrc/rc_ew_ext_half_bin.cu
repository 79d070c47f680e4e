// rc_ew_ext_half_bin.cu -- c = a o b for f16 / bf16 (see rc_ew_ext.cuh)
#include "rc_ew_ext.cuh"

namespace rc {

bool run_binary_half(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
#define RC_HALF_OP(OPCODE, W, FF)                                              \
    case OPCODE:                                                               \
        if (t == RC_F16) { ew_launch<W<h16, FF>>(dev, c, args); return true; }  \
        if (t == RC_BF16) { ew_launch<W<b16, FF>>(dev, c, args); return true; } \
        break;
    if (dtype_is_half(t)) {
        switch (op) {
            RC_HALF_OP(RC_ADD, FViaF32, FAdd) RC_HALF_OP(RC_SUB, FViaF32, FSub) RC_HALF_OP(RC_MUL, FViaF32, FMul)
            RC_HALF_OP(RC_DIV, FViaF32, FDiv) RC_HALF_OP(RC_REM, FViaF32, FRem)
            RC_HALF_OP(RC_MAXIMUM, FViaF32, FMaximum) RC_HALF_OP(RC_MINIMUM, FViaF32, FMinimum)
            RC_HALF_OP(RC_FLOOR_DIVIDE, FViaF32, FFloorDivide) RC_HALF_OP(RC_POW, FViaF32, FPow)
            RC_HALF_OP(RC_ATAN2, FViaF32, FAtan2) RC_HALF_OP(RC_COPYSIGN, FViaF32, FCopysign) RC_HALF_OP(RC_HYPOT, FViaF32, FHypot)
            RC_HALF_OP(RC_LOGADDEXP, FViaF32, FLogAddExp)
            RC_HALF_OP(RC_EQ, FViaF32Bool, FEq) RC_HALF_OP(RC_NE, FViaF32Bool, FNe) RC_HALF_OP(RC_LT, FViaF32Bool, FLt)
            RC_HALF_OP(RC_LE, FViaF32Bool, FLe) RC_HALF_OP(RC_GT, FViaF32Bool, FGt) RC_HALF_OP(RC_GE, FViaF32Bool, FGe)
            default: break;
        }
        return false;
    }
#undef RC_HALF_OP
    return false;
}

}  // namespace rc
