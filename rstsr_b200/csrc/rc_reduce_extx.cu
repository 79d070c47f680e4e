// rc_reduce_extx.cu -- sum / prod / mean (f16, bf16, c32, c64) and max / min (f16, bf16) on the reduction kernels of
// rc_reduce.cuh.  Half inputs are accumulated in f32 and rounded ONCE at the end (the reference accumulates in the
// element type, `acc + x` in half precision: our result is the more accurate one and lies within half-precision
// rounding of it; tests compare against an f64 sum with 2^-9 / 2^-6 relative tolerance for f16 / bf16).  Complex sums
// are componentwise; the complex mean divides by Complex::from(n) with the same division formula the operators use.
#include "rc_reduce.cuh"
#include "rc_types.cuh"

namespace rc {
namespace {

// half input, f32 state, half output; the second pass folds f32 states (PState)
template <class T, template <class> class PF>
struct PViaF32 {
    using TI = T; using S = float; using TO = T; using Second = PState<PViaF32<T, PF>>;
    static __device__ __forceinline__ S init() { return PF<float>::init(); }
    static __device__ __forceinline__ S pre(T x, int64_t) { return x.f(); }
    static __device__ __forceinline__ S comb(S a, S b) { return PF<float>::comb(a, b); }
    static __device__ __forceinline__ TO fin(S s, int64_t n) { return T(PF<float>::fin(s, n)); }
};

template <class T>
void reduce_half(rc_device *dev, rc_redop op, const CanonRed &c, const void *a, void *out, int64_t n) {
    switch (op) {
        case RC_SUM: reduce_typed<PViaF32<T, PSum>>(dev, c, a, out, n); return;
        case RC_PROD: reduce_typed<PViaF32<T, PProd>>(dev, c, a, out, n); return;
        case RC_MAX: reduce_typed<PViaF32<T, PMax>>(dev, c, a, out, n); return;
        case RC_MIN: reduce_typed<PViaF32<T, PMin>>(dev, c, a, out, n); return;
        case RC_MEAN: reduce_typed<PViaF32<T, PMean>>(dev, c, a, out, n); return;
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, "this reduction is not implemented for half types");
}

template <class T>
void reduce_cplx(rc_device *dev, rc_redop op, const CanonRed &c, const void *a, void *out, int64_t n) {
    switch (op) {
        case RC_SUM: reduce_typed<PSum<T>>(dev, c, a, out, n); return;
        case RC_PROD: reduce_typed<PProd<T>>(dev, c, a, out, n); return;
        case RC_MEAN: reduce_typed<PMean<T>>(dev, c, a, out, n); return;
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, "complex numbers have sum / prod / mean only (no ordering, ExtReal is not implemented for them)");
}

}  // namespace

void run_reduce_extx(rc_device *dev, rc_redop op, rc_dtype t, const CanonRed &cr, const void *a, void *out, int64_t n) {
    switch (t) {
        case RC_F16: reduce_half<h16>(dev, op, cr, a, out, n); return;
        case RC_BF16: reduce_half<b16>(dev, op, cr, a, out, n); return;
        case RC_C32: reduce_cplx<c32>(dev, op, cr, a, out, n); return;
        case RC_C64: reduce_cplx<c64>(dev, op, cr, a, out, n); return;
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "not an extended dtype");
}

}  // namespace rc
