// rc_ew_arith.cu -- Op{Add,Sub,Mul,Div,Rem}API + maximum/minimum/floor_divide for every numeric dtype
// (rstsr-core/src/feature_rayon/auto_impl/op_ternary_arithmetic.rs:3-56, op_ternary_common.rs).
#include "rc_dispatch.cuh"

namespace rc {

void run_binary_arith(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    switch (op) {
        case RC_ADD: switch (t) { RC_SWITCH_NUM(FAdd) default: break; } break;
        case RC_SUB: switch (t) { RC_SWITCH_NUM(FSub) default: break; } break;
        case RC_MUL: switch (t) { RC_SWITCH_NUM(FMul) default: break; } break;
        case RC_DIV: switch (t) { RC_SWITCH_NUM(FDiv) default: break; } break;
        case RC_REM: switch (t) { RC_SWITCH_NUM(FRem) default: break; } break;
        case RC_MAXIMUM: switch (t) { RC_SWITCH_NUM(FMaximum) default: break; } break;
        case RC_MINIMUM: switch (t) { RC_SWITCH_NUM(FMinimum) default: break; } break;
        case RC_FLOOR_DIVIDE: switch (t) { RC_SWITCH_NUM(FFloorDivide) default: break; } break;
        default: break;
    }
    unsupported("arithmetic op", t);
}

}  // namespace rc
