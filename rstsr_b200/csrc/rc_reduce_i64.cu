// rc_reduce_i64.cu -- typed instantiations of the reduction kernels (see rc_reduce.cuh).
#include "rc_reduce.cuh"

namespace rc {
void run_reduce_i64(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op<int64_t>(dev, op, cr, a, out, n);
}
void run_reduce_u64(rc_device *dev, rc_redop op, const CanonRed &cr, const void *a, void *out, int64_t n) {
    reduce_op<uint64_t>(dev, op, cr, a, out, n);
}
}
