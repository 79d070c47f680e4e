// rc_ew_unary.cu -- unary elementwise ops a = f(b)
// (rstsr-core/src/feature_rayon/auto_impl/op_binary_arithmetic.rs:94-113, op_binary_common.rs:10-239).
#include "rc_dispatch.cuh"

namespace rc {

void run_unary(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    switch (op) {
        case RC_NEG: switch (t) { RC_SWITCH_SIGNED_INT(FNeg) RC_SWITCH_FLOAT(FNeg) default: break; } break;
        case RC_NOT:
            if (t == RC_BOOL) { ew_launch<FNotBool>(dev, c, args); return; }
            switch (t) { RC_SWITCH_INT(FNot) default: break; }
            break;
        case RC_ABS: switch (t) { RC_SWITCH_NUM(FAbs) default: break; } break;
        case RC_SQUARE: switch (t) { RC_SWITCH_NUM(FSquare) default: break; } break;
        case RC_SIGN: switch (t) { RC_SWITCH_NUM(FSign) default: break; } break;
        case RC_CONJ:
        case RC_REAL: switch (t) { RC_SWITCH_NUM(FIdentity) default: break; } break;
        case RC_IMAG: switch (t) { RC_SWITCH_NUM(FZero) default: break; } break;
        case RC_SQRT: switch (t) { RC_SWITCH_FLOAT(FSqrt) default: break; } break;
        case RC_EXP: switch (t) { RC_SWITCH_FLOAT(FExp) default: break; } break;
        case RC_EXPM1: switch (t) { RC_SWITCH_FLOAT(FExpm1) default: break; } break;
        case RC_LOG: switch (t) { RC_SWITCH_FLOAT(FLog) default: break; } break;
        case RC_LOG2: switch (t) { RC_SWITCH_FLOAT(FLog2) default: break; } break;
        case RC_LOG10: switch (t) { RC_SWITCH_FLOAT(FLog10) default: break; } break;
        case RC_SIN: switch (t) { RC_SWITCH_FLOAT(FSin) default: break; } break;
        case RC_COS: switch (t) { RC_SWITCH_FLOAT(FCos) default: break; } break;
        case RC_TAN: switch (t) { RC_SWITCH_FLOAT(FTan) default: break; } break;
        case RC_ASIN: switch (t) { RC_SWITCH_FLOAT(FAsin) default: break; } break;
        case RC_ACOS: switch (t) { RC_SWITCH_FLOAT(FAcos) default: break; } break;
        case RC_ATAN: switch (t) { RC_SWITCH_FLOAT(FAtan) default: break; } break;
        case RC_SINH: switch (t) { RC_SWITCH_FLOAT(FSinh) default: break; } break;
        case RC_COSH: switch (t) { RC_SWITCH_FLOAT(FCosh) default: break; } break;
        case RC_TANH: switch (t) { RC_SWITCH_FLOAT(FTanh) default: break; } break;
        case RC_ASINH: switch (t) { RC_SWITCH_FLOAT(FAsinh) default: break; } break;
        case RC_ACOSH: switch (t) { RC_SWITCH_FLOAT(FAcosh) default: break; } break;
        case RC_ATANH: switch (t) { RC_SWITCH_FLOAT(FAtanh) default: break; } break;
        case RC_FLOOR: switch (t) { RC_SWITCH_FLOAT(FFloor) default: break; } break;
        case RC_CEIL: switch (t) { RC_SWITCH_FLOAT(FCeil) default: break; } break;
        case RC_ROUND: switch (t) { RC_SWITCH_FLOAT(FRound) default: break; } break;
        case RC_TRUNC: switch (t) { RC_SWITCH_FLOAT(FTrunc) default: break; } break;
        case RC_RECIPROCAL: switch (t) { RC_SWITCH_FLOAT(FRecip) default: break; } break;
        case RC_ISNAN: switch (t) { RC_SWITCH_FLOAT(FIsNan) default: break; } break;
        case RC_ISINF: switch (t) { RC_SWITCH_FLOAT(FIsInf) default: break; } break;
        case RC_ISFINITE: switch (t) { RC_SWITCH_FLOAT(FIsFinite) default: break; } break;
        case RC_SIGNBIT: switch (t) { RC_SWITCH_FLOAT(FSignBit) default: break; } break;
        default: break;
    }
    unsupported("unary op", t);
}

}  // namespace rc
