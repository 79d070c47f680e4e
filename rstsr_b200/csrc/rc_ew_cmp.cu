// rc_ew_cmp.cu -- comparisons with bool output (rstsr-core/src/feature_rayon/auto_impl/op_ternary_common.rs:89-94).
#include "rc_dispatch.cuh"

namespace rc {

void run_binary_cmp(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    switch (op) {
        case RC_EQ: switch (t) { RC_SWITCH_NUM(FEq) RC_CASE(RC_BOOL, uint8_t, FEq) default: break; } break;
        case RC_NE: switch (t) { RC_SWITCH_NUM(FNe) RC_CASE(RC_BOOL, uint8_t, FNe) default: break; } break;
        case RC_LT: switch (t) { RC_SWITCH_NUM(FLt) RC_CASE(RC_BOOL, uint8_t, FLt) default: break; } break;
        case RC_LE: switch (t) { RC_SWITCH_NUM(FLe) RC_CASE(RC_BOOL, uint8_t, FLe) default: break; } break;
        case RC_GT: switch (t) { RC_SWITCH_NUM(FGt) RC_CASE(RC_BOOL, uint8_t, FGt) default: break; } break;
        case RC_GE: switch (t) { RC_SWITCH_NUM(FGe) RC_CASE(RC_BOOL, uint8_t, FGe) default: break; } break;
        default: break;
    }
    unsupported("comparison", t);
}

}  // namespace rc
