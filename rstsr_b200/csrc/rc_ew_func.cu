// rc_ew_func.cu -- float binary functions pow/atan2/copysign/hypot/logaddexp/nextafter
// (rstsr-core/src/feature_rayon/auto_impl/op_ternary_common.rs:7-13 and the OpPowAPI block).
#include "rc_dispatch.cuh"

namespace rc {

void run_binary_func(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    switch (op) {
        case RC_POW: switch (t) { RC_SWITCH_FLOAT(FPow) default: break; } break;
        case RC_ATAN2: switch (t) { RC_SWITCH_FLOAT(FAtan2) default: break; } break;
        case RC_COPYSIGN: switch (t) { RC_SWITCH_FLOAT(FCopysign) default: break; } break;
        case RC_HYPOT: switch (t) { RC_SWITCH_FLOAT(FHypot) default: break; } break;
        case RC_LOGADDEXP: switch (t) { RC_SWITCH_FLOAT(FLogAddExp) default: break; } break;
        case RC_NEXTAFTER: switch (t) { RC_SWITCH_FLOAT(FNextAfter) default: break; } break;
        default: break;
    }
    unsupported("binary function", t);
}

// pow with a mixed exponent: t = base type; the exponent operand is i32 (float base) or u32 (integer base)
void run_binary_pow_mixed(rc_device *dev, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    switch (t) {
        RC_SWITCH_FLOAT(FPowi)
        RC_SWITCH_INT(FIPow)
        default: break;
    }
    unsupported("pow", t);
}

// elementwise isclose, bool output; args.params -> IsCloseParams
bool run_isclose_ext(rc_device *dev, rc_dtype t, const CanonEw &c, const EwArgs &args);  // rc_ew_ext_cplx.cu

void run_isclose(rc_device *dev, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    if (dtype_is_extended(t) && run_isclose_ext(dev, t, c, args)) return;
    switch (t) {
        RC_SWITCH_NUM(FIsClose)
        default: break;
    }
    unsupported("isclose", t);
}

}  // namespace rc
