// rc_reduce_bin.cu -- binary reductions on the reduce kernels of rc_reduce.cuh (two input streams, one pass):
//   rc_vecdot        c[m] = sum_s a[m, s] * b[m, s]     DeviceVecdotAPI::vecdot
//                    (rstsr-core/src/device_cpu_serial/linalg/vecdot.rs:4-29,
//                     rstsr-native-impl/src/cpu_serial/vecdot.rs:6-168; conj is the identity for real types)
//   rc_allclose_all  all(|a - b| <= atol + rtol * |b|)  OpAllCloseAPI::allclose_all
//                    (rstsr-core/src/device_cpu_serial/reduction.rs:660-683, rstsr-dtype-traits/src/isclose.rs:92-106)
// Both read every element of a and b once: HBM roofline, 2 x itemsize bytes per pair.
#include "rc_device.hpp"
#include "rc_layout.hpp"
#include "rc_reduce.cuh"

namespace rc {
namespace {

template <class T> struct PDot {
    static constexpr bool BINARY = true;
    using TI = T; using S = T; using TO = T; using Second = PSum<T>;
    static __device__ __forceinline__ S init() { return (T)0; }
    static __device__ __forceinline__ S pre2(T x, T y, const RedDesc &) {
        if constexpr (std::is_integral<T>::value) return (T)((uns<T>)x * (uns<T>)y); else return x * y;
    }
    static __device__ __forceinline__ S comb(S a, S b) { return PSum<T>::comb(a, b); }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};

// isclose with TE = f64: diff and |b| are formed in the element type, then cast (isclose.rs:100-105).
// inf vs inf gives |inf - inf| = NaN -> not close, as in the reference (NumPy says close).
template <class T> struct PClose {
    static constexpr bool BINARY = true;
    using TI = T; using S = uint8_t; using TO = uint8_t; using Second = PLogic<true>;
    static __device__ __forceinline__ S init() { return 1; }
    static __device__ __forceinline__ S pre2(T a, T b, const RedDesc &d) {
        double diff, abs_b;
        if constexpr (std::is_floating_point<T>::value) {
            const T df = a - b;  // rounded in the element type before the cast, like (self - other).abs()
            diff = (double)(df < (T)0 ? -df : df);
            abs_b = (double)(b < (T)0 ? -b : b);
            if (df != df) diff = (double)df;  // keep NaN
            if (b != b) abs_b = (double)b;
        } else if constexpr (std::is_signed<T>::value) {
            // ext_abs_diff: self >= other ? self - other : other - self (wrapping); ext_abs: wrapping abs
            const T df = a >= b ? (T)((uns<T>)a - (uns<T>)b) : (T)((uns<T>)b - (uns<T>)a);
            diff = (double)df;
            abs_b = (double)(b < 0 ? (T)((uns<T>)0 - (uns<T>)b) : b);
        } else {
            diff = (double)(a >= b ? (T)(a - b) : (T)(b - a));
            abs_b = (double)b;
        }
        bool ok = diff <= d.fp1 + d.fp0 * abs_b;  // atol + rtol * |b|
        if constexpr (std::is_floating_point<T>::value) ok = ok || (d.ip0 && a != a && b != b);
        return ok ? 1 : 0;
    }
    static __device__ __forceinline__ S comb(S a, S b) { return a & b; }
    static __device__ __forceinline__ TO fin(S s, int64_t) { return s; }
};

void check_dev_ptr(const void *p, const char *name) {
    if (!p) raise(RC_ERR_INVALID_VALUE, std::string("null device pointer: ") + name);
}

}  // namespace
// half / complex element types (rc_reduce_extx.cu)
void run_vecdot_extx(rc_device *dev, rc_dtype t, const CanonRed &cr, const void *a, const void *b, void *c, int64_t n);
void run_allclose_extx(rc_device *dev, rc_dtype t, const CanonRed &cr, const void *a, const void *b, void *out, int64_t n,
                       double rtol, double atol, int equal_nan);
namespace {

template <template <class> class P>
void dispatch(rc_device *dev, rc_dtype t, const CanonRed &cr, const void *a, const void *b, void *out, int64_t n,
              double fp0, double fp1, int ip0) {
    if (dtype_is_extended(t)) {
        if (std::is_same<P<float>, PDot<float>>::value) run_vecdot_extx(dev, t, cr, a, b, out, n);
        else run_allclose_extx(dev, t, cr, a, b, out, n, fp0, fp1, ip0);
        return;
    }
    switch (t) {
        case RC_F64: reduce_typed<P<double>>(dev, cr, a, out, n, b, fp0, fp1, ip0); return;
        case RC_F32: reduce_typed<P<float>>(dev, cr, a, out, n, b, fp0, fp1, ip0); return;
        case RC_I64: reduce_typed<P<int64_t>>(dev, cr, a, out, n, b, fp0, fp1, ip0); return;
        case RC_U64: reduce_typed<P<uint64_t>>(dev, cr, a, out, n, b, fp0, fp1, ip0); return;
        case RC_I32: reduce_typed<P<int32_t>>(dev, cr, a, out, n, b, fp0, fp1, ip0); return;
        case RC_U32: reduce_typed<P<uint32_t>>(dev, cr, a, out, n, b, fp0, fp1, ip0); return;
        default: raise(RC_ERR_UNIMPLEMENTED, "binary reductions are implemented for f32, f64, i32, u32, i64, u64");
    }
}

// broadcast_layout_to_first (rstsr-common/src/layout/broadcast.rs): l must broadcast to lc's shape exactly
Layout broadcast_to(const Layout &lc, const Layout &l, rc_order order) {
    Layout oc, ol;
    broadcast_layouts(lc, l, order, &oc, &ol);
    RC_CHECK(oc.shape == lc.shape, RC_ERR_INVALID_LAYOUT, "layout of c seems not broadcasted from a or b after axis sum");
    return ol;
}

}  // namespace
}  // namespace rc

using namespace rc;

extern "C" {

int rc_vecdot(rc_device *dev, rc_dtype t, void *c, const rc_layout *lc_, const void *a, const rc_layout *la_,
              const void *b, const rc_layout *lb_, const int64_t *axes_a, const int64_t *axes_b, int naxes) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout lc = from_c(lc_), la = from_c(la_), lb = from_c(lb_);
        std::vector<int> ax_a = normalize_axes(axes_a, naxes, la.ndim());
        std::vector<int> ax_b = normalize_axes(axes_b, naxes, lb.ndim());
        Layout las, lam, lbs, lbm;
        split_axes(la, ax_a, &las, &lam, nullptr);
        split_axes(lb, ax_b, &lbs, &lbm, nullptr);
        RC_CHECK(las.shape == lbs.shape, RC_ERR_INVALID_LAYOUT,
                 "the dimensions of a and b along the contracted axis should be the same");
        Layout lam_b = broadcast_to(lc, lam, dev->order);
        Layout lbm_b = broadcast_to(lc, lbm, dev->order);
        if (lc.size() == 0) return;
        check_dev_ptr(c, "c");
        if (las.size() != 0) { check_dev_ptr(a, "a"); check_dev_ptr(b, "b"); }
        CanonRed cr = canon_reduce_binary(lam_b, lbm_b, lc, las, lbs, la.offset, lb.offset);
        std::lock_guard<std::mutex> lock(dev->ws_mu);
        dispatch<PDot>(dev, t, cr, a, b, c, las.size(), 0.0, 0.0, 0);
    });
}

int rc_allclose_all(rc_device *dev, rc_dtype t, const void *a, const rc_layout *la_, const void *b, const rc_layout *lb_,
                    double rtol, double atol, int equal_nan, int *result) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(result, RC_ERR_INVALID_VALUE, "null result");
        Layout la = from_c(la_), lb = from_c(lb_);
        RC_CHECK(la.shape == lb.shape, RC_ERR_INVALID_LAYOUT, "allclose_all takes layouts already broadcast to one shape");
        RC_CHECK(la.size() != 0 && lb.size() != 0, RC_ERR_INVALID_VALUE, "zero-size array is not supported for allclose");
        check_dev_ptr(a, "a"); check_dev_ptr(b, "b");
        Layout kept;  // 0-d: everything is reduced
        CanonRed cr = canon_reduce_binary(kept, kept, kept, la, lb, la.offset, lb.offset);
        std::lock_guard<std::mutex> slot_lock(dev->slot_mu);
        void *host_slot = nullptr;
        void *slot = scalar_slot(dev, &host_slot);
        {
            std::lock_guard<std::mutex> lock(dev->ws_mu);
            dispatch<PClose>(dev, t, cr, a, b, slot, la.size(), rtol, atol, equal_nan ? 1 : 0);
        }
        RC_CUDA(cudaStreamSynchronize(dev->stream));
        const uint8_t host = *static_cast<const volatile uint8_t *>(host_slot);
        *result = host ? 1 : 0;
    });
}

}  // extern "C"
