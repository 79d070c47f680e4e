// rc_reduce_i8.cu -- sum / prod / max / min and argmin / argmax / count_nonzero for i8, u8 (see rc_reduce.cuh).
// Sums and products wrap in the element type, as the reference's release-mode Rust does.
#include "rc_reduce.cuh"

namespace rc {
void run_reduce_i8(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    if (op >= RC_VAR) reduce_op_ext<int8_t>(dev, op, cr, a, out, n); else reduce_op<int8_t>(dev, op, cr, a, out, n);
}
void run_reduce_u8(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    if (op >= RC_VAR) reduce_op_ext<uint8_t>(dev, op, cr, a, out, n); else reduce_op<uint8_t>(dev, op, cr, a, out, n);
}
}
