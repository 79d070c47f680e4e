// rc_reduce_i32.cu -- typed instantiations of the reduction kernels (see rc_reduce.cuh).
#include "rc_reduce.cuh"

namespace rc {
void run_reduce_i32(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op<int32_t>(dev, op, cr, a, out, n);
}
void run_reduce_u32(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op<uint32_t>(dev, op, cr, a, out, n);
}
}
