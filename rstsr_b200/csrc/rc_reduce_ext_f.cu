// rc_reduce_ext_f.cu -- var / std / l2_norm / argmin / argmax / count_nonzero for f32, f64 (see rc_reduce.cuh).
#include "rc_reduce.cuh"

namespace rc {
void run_reduce_ext_f64(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op_ext<double>(dev, op, cr, a, out, n);
}
void run_reduce_ext_f32(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op_ext<float>(dev, op, cr, a, out, n);
}
}
