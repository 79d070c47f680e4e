// rc_reduce_ext_i.cu -- argmin / argmax / count_nonzero for integers; all / any / count for bool (see rc_reduce.cuh).
#include "rc_reduce.cuh"

namespace rc {
void run_reduce_ext_i64(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op_ext<int64_t>(dev, op, cr, a, out, n);
}
void run_reduce_ext_u64(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op_ext<uint64_t>(dev, op, cr, a, out, n);
}
void run_reduce_ext_i32(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op_ext<int32_t>(dev, op, cr, a, out, n);
}
void run_reduce_ext_u32(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op_ext<uint32_t>(dev, op, cr, a, out, n);
}
void run_reduce_bool(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    switch (op) {
        case RC_ALL: reduce_typed<PLogic<true>>(dev, cr, a, out, n); return;
        case RC_ANY: reduce_typed<PLogic<false>>(dev, cr, a, out, n); return;
        case RC_COUNT_NONZERO: reduce_typed<PCount<uint8_t>>(dev, cr, a, out, n); return;  // also OpSumBoolAPI
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, "this reduction is not implemented for bool");
}
}
