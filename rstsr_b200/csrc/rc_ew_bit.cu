// rc_ew_bit.cu -- Op{BitOr,BitAnd,BitXor,Shl,Shr}API for integer dtypes (and bool for | & ^)
// (rstsr-core/src/feature_rayon/auto_impl/op_ternary_arithmetic.rs:9-14).
#include "rc_dispatch.cuh"

namespace rc {

void run_binary_bit(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    switch (op) {
        case RC_BITOR: switch (t) { RC_SWITCH_INT(FBitOr) RC_CASE(RC_BOOL, uint8_t, FBitOr) default: break; } break;
        case RC_BITAND: switch (t) { RC_SWITCH_INT(FBitAnd) RC_CASE(RC_BOOL, uint8_t, FBitAnd) default: break; } break;
        case RC_BITXOR: switch (t) { RC_SWITCH_INT(FBitXor) RC_CASE(RC_BOOL, uint8_t, FBitXor) default: break; } break;
        case RC_SHL: switch (t) { RC_SWITCH_INT(FShl) default: break; } break;
        case RC_SHR: switch (t) { RC_SWITCH_INT(FShr) default: break; } break;
        default: break;
    }
    unsupported("bit op", t);
}

}  // namespace rc
