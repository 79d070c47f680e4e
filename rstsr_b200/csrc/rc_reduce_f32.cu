// rc_reduce_f32.cu -- typed instantiations of the reduction kernels (see rc_reduce.cuh).
#include "rc_reduce.cuh"

namespace rc {
void run_reduce_f32(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op<float>(dev, op, cr, a, out, n);
}
}
