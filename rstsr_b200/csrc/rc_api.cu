// rc_api.cu -- the extern "C" boundary (include/rstsr_cuda.h): device handles, raw storage, layout helpers and
// the per-trait compute entry points.  Each entry point canonicalises its layouts on the host and enqueues
// the kernels on the handle's stream; nothing here falls back to the CPU.
#include <algorithm>
#include <cmath>

#include "rc_canon.hpp"
#include "rc_device.hpp"
#include "rc_elementwise.cuh"
#include "rc_layout.hpp"
#include "rc_ops.hpp"

namespace rc {

const std::string &last_error_ref();
bool host_free_numa(void *ptr);  // rc_api_ex.cu

void *workspace(rc_device *d, size_t nbytes) {
    if (nbytes <= d->ws_bytes) return d->ws;
    size_t want = std::max(nbytes, std::max<size_t>(2 * d->ws_bytes, 1 << 20));
    if (d->ws) RC_CUDA(cudaFreeAsync(d->ws, d->stream));  // stream-ordered: earlier kernels finish first
    d->ws = nullptr;
    d->ws_bytes = 0;
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, want, d->stream);
    if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("workspace allocation failed: ") + cudaGetErrorString(e));
    d->ws = p;
    d->ws_bytes = want;
    return p;
}

void upload_small(rc_device *d, void *dst_dev, const void *src, size_t nbytes) {
    if (nbytes == 0) return;
    if (nbytes > rc_device::STAGE_BYTES) {
        RC_CUDA(cudaMemcpyAsync(dst_dev, src, nbytes, cudaMemcpyHostToDevice, d->stream));
        return;
    }
    std::lock_guard<std::mutex> lock(d->stage_mu);
    const int s = d->stage_next;
    d->stage_next = (s + 1) % rc_device::STAGE_SLOTS;
    if (!d->stage_buf[s]) {
        RC_CUDA(cudaHostAlloc(&d->stage_buf[s], rc_device::STAGE_BYTES, cudaHostAllocPortable));
        RC_CUDA(cudaEventCreateWithFlags(&d->stage_ev[s], cudaEventDisableTiming));
    } else {
        RC_CUDA(cudaEventSynchronize(d->stage_ev[s]));  // the copy that last used this slot has left it
    }
    std::memcpy(d->stage_buf[s], src, nbytes);
    RC_CUDA(cudaMemcpyAsync(dst_dev, d->stage_buf[s], nbytes, cudaMemcpyHostToDevice, d->stream));
    RC_CUDA(cudaEventRecord(d->stage_ev[s], d->stream));
}

void *scalar_slot(rc_device *d, void **host) {
    if (!d->slot_host) {
        void *h = nullptr, *p = nullptr;
        RC_CUDA(cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable));
        cudaError_t e = cudaHostGetDevicePointer(&p, h, 0);
        if (e != cudaSuccess) { cudaFreeHost(h); raise(RC_ERR_DEVICE, std::string("cudaHostGetDevicePointer: ") + cudaGetErrorString(e)); }
        d->slot_host = h;
        d->slot_dev = p;
    }
    *host = d->slot_host;
    return d->slot_dev;
}

void run_reduce_f64(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_f32(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_i64(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_u64(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_i32(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_u32(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_ext_f64(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_ext_f32(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_ext_i64(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_ext_u64(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_ext_i32(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_ext_u32(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_bool(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_i8(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_u8(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_i16(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);
void run_reduce_u16(rc_device *, rc_redop, const CanonRed &, const void *, void *, int64_t);

rc_dtype redop_out_dtype(rc_redop op, rc_dtype t) {
    // var / std / l2_norm of a complex tensor are real (TOut = T::Real, auto_impl/reduction.rs:213,267,323)
    if (dtype_is_complex(t) && (op == RC_VAR || op == RC_STD || op == RC_L2_NORM)) return t == RC_C32 ? RC_F32 : RC_F64;
    switch (op) {
        case RC_ARGMIN: case RC_ARGMAX: case RC_COUNT_NONZERO: return RC_U64;
        case RC_ALL: case RC_ANY: return RC_BOOL;
        default: return t;
    }
}

void run_reduce(rc_device *dev, rc_redop op, rc_dtype t, const CanonRed &cr, const void *a, void *out,
                int64_t mean_count) {
    if (dtype_is_extended(t)) { run_reduce_extx(dev, op, t, cr, a, out, mean_count); return; }
    switch (t) {  // narrow integers: base and "next" ops live in one TU per width
        case RC_I8: run_reduce_i8(dev, op, cr, a, out, mean_count); return;
        case RC_U8: run_reduce_u8(dev, op, cr, a, out, mean_count); return;
        case RC_I16: run_reduce_i16(dev, op, cr, a, out, mean_count); return;
        case RC_U16: run_reduce_u16(dev, op, cr, a, out, mean_count); return;
        default: break;
    }
    if (op >= RC_VAR) {  // the "next" reductions
        switch (t) {
            case RC_F64: run_reduce_ext_f64(dev, op, cr, a, out, mean_count); return;
            case RC_F32: run_reduce_ext_f32(dev, op, cr, a, out, mean_count); return;
            case RC_I64: run_reduce_ext_i64(dev, op, cr, a, out, mean_count); return;
            case RC_U64: run_reduce_ext_u64(dev, op, cr, a, out, mean_count); return;
            case RC_I32: run_reduce_ext_i32(dev, op, cr, a, out, mean_count); return;
            case RC_U32: run_reduce_ext_u32(dev, op, cr, a, out, mean_count); return;
            case RC_BOOL: run_reduce_bool(dev, op, cr, a, out, mean_count); return;
            default: break;
        }
        raise(RC_ERR_UNIMPLEMENTED, std::string("reduction is not implemented for dtype ") + dtype_name(t));
    }
    switch (t) {
        case RC_F64: run_reduce_f64(dev, op, cr, a, out, mean_count); return;
        case RC_F32: run_reduce_f32(dev, op, cr, a, out, mean_count); return;
        case RC_I64: run_reduce_i64(dev, op, cr, a, out, mean_count); return;
        case RC_U64: run_reduce_u64(dev, op, cr, a, out, mean_count); return;
        case RC_I32: run_reduce_i32(dev, op, cr, a, out, mean_count); return;
        case RC_U32: run_reduce_u32(dev, op, cr, a, out, mean_count); return;
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, std::string("reduction is not implemented for dtype ") + dtype_name(t));
}

namespace {

void check_ptr(const void *p, const char *name) {
    RC_CHECK(p != nullptr, RC_ERR_INVALID_VALUE, std::string("null pointer: ") + name);
}

// Rust `as` from a host scalar of dtype `tf` to dtype `tc` (rstsr-dtype-traits/src/promotion.rs); 8 bytes out
template <class TOut, class TIn>
TOut host_cast(TIn v) {
    if constexpr (std::is_floating_point<TIn>::value && std::is_integral<TOut>::value) {
        if (v != v) return (TOut)0;
        const long double lo = (long double)std::numeric_limits<TOut>::min();
        const long double hi = (long double)std::numeric_limits<TOut>::max();
        long double x = std::trunc((long double)v);
        if (x <= lo) return std::numeric_limits<TOut>::min();
        if (x >= hi) return std::numeric_limits<TOut>::max();
        return (TOut)x;
    } else {
        return (TOut)v;
    }
}

template <class TIn>
void host_cast_to(rc_dtype tc, TIn v, bool in_bool, void *out8) {
    std::memset(out8, 0, 8);
    if (tc == RC_BOOL) { uint8_t b = (v != (TIn)0) ? 1 : 0; std::memcpy(out8, &b, 1); return; }
    if (in_bool) v = (TIn)(v != (TIn)0 ? 1 : 0);
#define RC_HC(DT, CT) case DT: { CT r = host_cast<CT, TIn>(v); std::memcpy(out8, &r, sizeof(CT)); return; }
    switch (tc) {
        RC_HC(RC_I8, int8_t) RC_HC(RC_I16, int16_t) RC_HC(RC_I32, int32_t) RC_HC(RC_I64, int64_t)
        RC_HC(RC_U8, uint8_t) RC_HC(RC_U16, uint16_t) RC_HC(RC_U32, uint32_t) RC_HC(RC_U64, uint64_t)
        RC_HC(RC_F32, float) RC_HC(RC_F64, double)
        default: break;
    }
#undef RC_HC
    raise(RC_ERR_INVALID_VALUE, "unknown dtype");
}

void host_scalar_cast_prim(rc_dtype tc, rc_dtype tf, const void *src, void *out8);

// out16: 16 bytes (c64 is the widest element)
void host_scalar_cast(rc_dtype tc, rc_dtype tf, const void *src, void *out16) {
    std::memset(out16, 0, 16);
    if (!dtype_is_extended(tc) && !dtype_is_extended(tf)) { host_scalar_cast_prim(tc, tf, src, out16); return; }
    if (tc == tf) { std::memcpy(out16, src, dtype_size(tc)); return; }
    double re = 0.0, im = 0.0;
    if (dtype_is_extended(tf)) host_from_ext(tf, src, &re, &im);
    else { unsigned char d8[8]; host_scalar_cast_prim(RC_F64, tf, src, d8); std::memcpy(&re, d8, 8); }
    if (dtype_is_extended(tc)) { host_to_ext(tc, re, im, out16); return; }
    RC_CHECK(!dtype_is_complex(tf), RC_ERR_UNIMPLEMENTED, "complex scalars do not cast to real types");
    host_scalar_cast_prim(tc, RC_F64, &re, out16);
}

void host_scalar_cast_prim(rc_dtype tc, rc_dtype tf, const void *src, void *out8) {
#define RC_HS(DT, CT, INB) case DT: { CT v; std::memcpy(&v, src, sizeof(CT)); host_cast_to<CT>(tc, v, INB, out8); return; }
    switch (tf) {
        RC_HS(RC_BOOL, uint8_t, true)
        RC_HS(RC_I8, int8_t, false) RC_HS(RC_I16, int16_t, false) RC_HS(RC_I32, int32_t, false) RC_HS(RC_I64, int64_t, false)
        RC_HS(RC_U8, uint8_t, false) RC_HS(RC_U16, uint16_t, false) RC_HS(RC_U32, uint32_t, false) RC_HS(RC_U64, uint64_t, false)
        RC_HS(RC_F32, float, false) RC_HS(RC_F64, double, false)
    }
#undef RC_HS
    raise(RC_ERR_INVALID_VALUE, "unknown dtype");
}

bool is_cmp(rc_binop op) { return op >= RC_EQ && op <= RC_GE; }
bool is_bit(rc_binop op) { return op >= RC_BITOR && op <= RC_SHR; }
bool is_func(rc_binop op) { return op >= RC_POW && op <= RC_NEXTAFTER; }
bool is_predicate(rc_unop op) { return op >= RC_ISNAN && op <= RC_SIGNBIT; }

void dispatch_binary(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args) {
    if (c.empty) return;
    if (dtype_is_extended(t)) {
        if (run_binary_ext(dev, op, t, c, args)) return;
        raise(RC_ERR_UNIMPLEMENTED, std::string("this binary op is not implemented for dtype ") + dtype_name(t));
    }
    if (is_cmp(op)) run_binary_cmp(dev, op, t, c, args);
    else if (is_bit(op)) run_binary_bit(dev, op, t, c, args);
    else if (is_func(op)) run_binary_func(dev, op, t, c, args);
    else run_binary_arith(dev, op, t, c, args);
}

std::vector<int> all_axes(int ndim) {
    std::vector<int> r(ndim);
    for (int i = 0; i < ndim; ++i) r[i] = i;
    return r;
}

void reduce_into_nolock(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const Layout &la,
                        const std::vector<int> &axes, void *out, const Layout &lo) {
    if (op == RC_MAX || op == RC_MIN)
        RC_CHECK(la.size() != 0, RC_ERR_INVALID_VALUE,
                 op == RC_MAX ? "zero-size array is not supported for max" : "zero-size array is not supported for min");
    if (op == RC_MEAN)
        RC_CHECK(dtype_is_float(t) || dtype_is_extended(t), RC_ERR_UNIMPLEMENTED, "mean requires a floating-point dtype");
    if (op == RC_VAR || op == RC_STD || op == RC_L2_NORM)
        RC_CHECK(dtype_is_float(t) || dtype_is_extended(t), RC_ERR_UNIMPLEMENTED, "var / std / l2_norm require a floating-point dtype");
    if (op == RC_ALL || op == RC_ANY) RC_CHECK(t == RC_BOOL, RC_ERR_UNIMPLEMENTED, "all / any take a bool tensor");
    const bool arg = (op == RC_ARGMIN || op == RC_ARGMAX);
    if (arg)  // reduce_all_unraveled_arg_cpu_serial: "empty sequence is not allowed for reduce_arg."
        RC_CHECK(la.size() != 0, RC_ERR_INVALID_LAYOUT, "empty sequence is not allowed for reduce_arg.");
    Layout l_axes;
    split_axes(la, axes, &l_axes, nullptr, nullptr);  // validates both halves like the reference
    CanonRed cr = canon_reduce(la, axes, lo, /*keep_order=*/arg);
    run_reduce(dev, op, t, cr, a, out, l_axes.size());
}

void reduce_into(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const Layout &la, const std::vector<int> &axes,
                 void *out, const Layout &lo) {
    std::lock_guard<std::mutex> lock(dev->ws_mu);
    dev->preq = rc_device::PartialReq();
    reduce_into_nolock(dev, op, t, a, la, axes, out, lo);
}

}  // namespace

void cast_host_scalar(rc_dtype tc, rc_dtype tf, const void *src, void *out8) { host_scalar_cast(tc, tf, src, out8); }

// for rc_comm.cu (sharded reductions): the caller holds dev->ws_mu and has set dev->preq
void reduce_local(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const Layout &la, const std::vector<int> &axes,
                  void *out, const Layout &lo) {
    reduce_into_nolock(dev, op, t, a, la, axes, out, lo);
}

}  // namespace rc

using namespace rc;

extern "C" {

const char *rc_last_error(void) { return last_error_ref().c_str(); }
const char *rc_version(void) { return "rstsr-cuda 0.1.0 sm_100a"; }

int rc_device_count(int *count) {
    return guard([&] {
        RC_CHECK(count != nullptr, RC_ERR_INVALID_VALUE, "null count");
        RC_CUDA(cudaGetDeviceCount(count));
    });
}

static int device_create(int ordinal, rc_order order, void *stream, bool borrow, rc_device **out) {
    return guard([&] {
        RC_CHECK(out != nullptr, RC_ERR_INVALID_VALUE, "null out");
        RC_CHECK(order == RC_ROW_MAJOR || order == RC_COL_MAJOR, RC_ERR_INVALID_VALUE, "invalid order");
        int n = 0;
        RC_CUDA(cudaGetDeviceCount(&n));
        RC_CHECK(ordinal >= 0 && ordinal < n, RC_ERR_DEVICE, "CUDA device ordinal out of range (no CPU fallback)");
        RC_CUDA(cudaSetDevice(ordinal));
        std::unique_ptr<rc_device> d(new rc_device());
        d->ordinal = ordinal;
        d->order = order;
        if (borrow) {
            d->stream = static_cast<cudaStream_t>(stream);
            d->own_stream = false;
        } else {
            RC_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
            d->own_stream = true;
        }
        int sms = 0;
        RC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ordinal));
        d->sm_count = sms > 0 ? sms : 148;
        // keep freed blocks cached in the default pool: rc_malloc/rc_free are on the hot path of `&a + &b`
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ordinal) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        *out = d.release();
    });
}

int rc_device_create(int ordinal, rc_order order, rc_device **out) { return device_create(ordinal, order, nullptr, false, out); }
int rc_device_create_on_stream(int ordinal, rc_order order, void *cuda_stream, rc_device **out) {
    return device_create(ordinal, order, cuda_stream, true, out);
}

int rc_device_destroy(rc_device *dev) {
    return guard([&] {
        if (!dev) return;
        cudaSetDevice(dev->ordinal);
        cudaStreamSynchronize(dev->stream);
        if (dev->ws) cudaFree(dev->ws);
        if (dev->slot_host) cudaFreeHost(dev->slot_host);
        for (int i = 0; i < rc_device::STAGE_SLOTS; ++i) {
            if (dev->stage_buf[i]) cudaFreeHost(dev->stage_buf[i]);
            if (dev->stage_ev[i]) cudaEventDestroy(dev->stage_ev[i]);
        }
        if (dev->own_stream) cudaStreamDestroy(dev->stream);
        delete dev;
    });
}

int rc_device_default_order(const rc_device *dev, rc_order *out) {
    return guard([&] { RC_CHECK(dev && out, RC_ERR_INVALID_VALUE, "null argument"); *out = dev->order; });
}
int rc_device_set_default_order(rc_device *dev, rc_order order) {
    return guard([&] {
        RC_CHECK(dev, RC_ERR_INVALID_VALUE, "null device");
        RC_CHECK(order == RC_ROW_MAJOR || order == RC_COL_MAJOR, RC_ERR_INVALID_VALUE, "invalid order");
        dev->order = order;
    });
}
int rc_device_same_device(const rc_device *a, const rc_device *b, int *same) {
    return guard([&] {
        RC_CHECK(a && b && same, RC_ERR_INVALID_VALUE, "null argument");
        *same = (a->ordinal == b->ordinal && a->order == b->order && a->stream == b->stream) ? 1 : 0;
    });
}
int rc_device_ordinal(const rc_device *dev, int *ordinal) {
    return guard([&] { RC_CHECK(dev && ordinal, RC_ERR_INVALID_VALUE, "null argument"); *ordinal = dev->ordinal; });
}
int rc_device_stream(const rc_device *dev, void **s) {
    return guard([&] { RC_CHECK(dev && s, RC_ERR_INVALID_VALUE, "null argument"); *s = (void *)dev->stream; });
}
int rc_device_synchronize(rc_device *dev) {
    return guard([&] { DeviceGuard g(dev); RC_CUDA(cudaStreamSynchronize(dev->stream)); });
}
int rc_device_launch_count(const rc_device *dev, uint64_t *count) {
    return guard([&] { RC_CHECK(dev && count, RC_ERR_INVALID_VALUE, "null argument"); *count = dev->launches.load(); });
}

/* ---------------- raw storage ---------------- */
int rc_malloc(rc_device *dev, size_t nbytes, void **out) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(out != nullptr, RC_ERR_INVALID_VALUE, "null out");
        *out = nullptr;
        if (nbytes == 0) nbytes = 1;  // a zero-length Vec still has an identity
        cudaError_t e = cudaMallocAsync(out, nbytes, dev->stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            raise(RC_ERR_MEMORY, std::string("cudaMallocAsync(") + std::to_string(nbytes) + "): " + cudaGetErrorString(e));
        }
    });
}
int rc_free(rc_device *dev, void *ptr) {
    return guard([&] {
        DeviceGuard g(dev);
        if (ptr) RC_CUDA(cudaFreeAsync(ptr, dev->stream));
    });
}
int rc_memcpy_h2d(rc_device *dev, void *dst, const void *src, size_t nbytes) {
    return guard([&] {
        DeviceGuard g(dev);
        if (nbytes == 0) return;
        check_ptr(dst, "dst"); check_ptr(src, "src");
        RC_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, dev->stream));
    });
}
int rc_memcpy_d2h(rc_device *dev, void *dst, const void *src, size_t nbytes) {
    return guard([&] {
        DeviceGuard g(dev);
        if (nbytes != 0) {
            check_ptr(dst, "dst"); check_ptr(src, "src");
            RC_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, dev->stream));
        }
        RC_CUDA(cudaStreamSynchronize(dev->stream));
    });
}
int rc_memcpy_d2d(rc_device *dev, void *dst, const void *src, size_t nbytes) {
    return guard([&] {
        DeviceGuard g(dev);
        if (nbytes == 0) return;
        check_ptr(dst, "dst"); check_ptr(src, "src");
        RC_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, dev->stream));
    });
}
int rc_memcpy_h2d_async(rc_device *dev, void *dst, const void *src, size_t nbytes) {
    return rc_memcpy_h2d(dev, dst, src, nbytes);  // already stream-ordered and non-blocking for pinned memory
}
int rc_memcpy_d2h_async(rc_device *dev, void *dst, const void *src, size_t nbytes) {
    return guard([&] {
        DeviceGuard g(dev);
        if (nbytes == 0) return;
        check_ptr(dst, "dst"); check_ptr(src, "src");
        RC_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, dev->stream));
    });
}
int rc_memcpy2d_d2h_async(rc_device *dev, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                          size_t height) {
    return guard([&] {
        DeviceGuard g(dev);
        if (width == 0 || height == 0) return;
        check_ptr(dst, "dst"); check_ptr(src, "src");
        RC_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, dev->stream));
    });
}
int rc_memcpy2d_h2d_async(rc_device *dev, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width,
                          size_t height) {
    return guard([&] {
        DeviceGuard g(dev);
        if (width == 0 || height == 0) return;
        check_ptr(dst, "dst"); check_ptr(src, "src");
        RC_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, dev->stream));
    });
}
int rc_device_wait(rc_device *waiter, rc_device *signaler) {
    return guard([&] {
        RC_CHECK(waiter && signaler, RC_ERR_INVALID_VALUE, "null device handle");
        RC_CHECK(waiter->ordinal == signaler->ordinal, RC_ERR_DEVICE_MISMATCH, "handles belong to different GPUs");
        DeviceGuard g(waiter);
        if (waiter->stream == signaler->stream) return;
        cudaEvent_t ev;
        RC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, signaler->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(waiter->stream, ev, 0);
        cudaEventDestroy(ev);  // released once the wait has been satisfied
        if (e != cudaSuccess) raise(RC_ERR_DEVICE, std::string("rc_device_wait: ") + cudaGetErrorString(e));
    });
}
int rc_memcpy_peer(rc_device *dst_dev, void *dst, rc_device *src_dev, const void *src, size_t nbytes) {
    return guard([&] {
        RC_CHECK(dst_dev && src_dev, RC_ERR_INVALID_VALUE, "null device handle");
        if (nbytes == 0) return;
        check_ptr(dst, "dst"); check_ptr(src, "src");
        // order: [work already on src stream] -> copy (on dst stream) -> [later work on either stream]
        cudaEvent_t ready = nullptr, done = nullptr;
        {
            DeviceGuard gs(src_dev);
            RC_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
            cudaError_t e = cudaEventRecord(ready, src_dev->stream);
            if (e != cudaSuccess) { cudaEventDestroy(ready); raise(RC_ERR_DEVICE, std::string("rc_memcpy_peer: ") + cudaGetErrorString(e)); }
        }
        cudaError_t e;
        {
            DeviceGuard gd(dst_dev);
            e = cudaStreamWaitEvent(dst_dev->stream, ready, 0);
            if (e == cudaSuccess)
                e = cudaMemcpyPeerAsync(dst, dst_dev->ordinal, src, src_dev->ordinal, nbytes, dst_dev->stream);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(done, dst_dev->stream);
        }
        if (e == cudaSuccess && done) {
            DeviceGuard gs(src_dev);
            e = cudaStreamWaitEvent(src_dev->stream, done, 0);  // a later rc_free of `src` is ordered after the copy
        }
        cudaEventDestroy(ready);
        if (done) cudaEventDestroy(done);
        if (e != cudaSuccess) raise(RC_ERR_DEVICE, std::string("rc_memcpy_peer: ") + cudaGetErrorString(e));
    });
}
int rc_memset(rc_device *dev, void *dst, int byte, size_t nbytes) {
    return guard([&] {
        DeviceGuard g(dev);
        if (nbytes == 0) return;
        check_ptr(dst, "dst");
        RC_CUDA(cudaMemsetAsync(dst, byte, nbytes, dev->stream));
    });
}
int rc_get_index(rc_device *dev, rc_dtype t, const void *a, int64_t index, void *host_out) {
    return guard([&] {
        DeviceGuard g(dev);
        check_ptr(a, "a"); check_ptr(host_out, "host_out");
        RC_CHECK(index >= 0, RC_ERR_INDEX, "negative index");
        size_t sz = dtype_size(t);
        RC_CUDA(cudaMemcpyAsync(host_out, static_cast<const char *>(a) + index * sz, sz, cudaMemcpyDeviceToHost, dev->stream));
        RC_CUDA(cudaStreamSynchronize(dev->stream));
    });
}
int rc_set_index(rc_device *dev, rc_dtype t, void *a, int64_t index, const void *host_value) {
    return guard([&] {
        DeviceGuard g(dev);
        check_ptr(a, "a"); check_ptr(host_value, "host_value");
        RC_CHECK(index >= 0, RC_ERR_INDEX, "negative index");
        size_t sz = dtype_size(t);
        RC_CUDA(cudaMemcpyAsync(static_cast<char *>(a) + index * sz, host_value, sz, cudaMemcpyHostToDevice, dev->stream));
        RC_CUDA(cudaStreamSynchronize(dev->stream));  // host_value may be a temporary
    });
}
int rc_host_alloc(size_t nbytes, void **out) {
    return guard([&] {
        RC_CHECK(out != nullptr, RC_ERR_INVALID_VALUE, "null out");
        cudaError_t e = cudaMallocHost(out, nbytes ? nbytes : 1);
        if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    });
}
int rc_host_free(void *ptr) {
    return guard([&] {
        if (!ptr) return;
        if (host_free_numa(ptr)) return;  // rc_host_alloc_on_node block (rc_api_ex.cu)
        RC_CUDA(cudaFreeHost(ptr));
    });
}
size_t rc_dtype_size(rc_dtype t) {
    size_t s = 0;
    guard([&] { s = dtype_size(t); });
    return s;
}

/* ---------------- layout helpers ---------------- */
int rc_layout_check(const rc_layout *l) { return guard([&] { check_layout(from_c(l)); }); }
int rc_layout_bounds_index(const rc_layout *l, int64_t *mn, int64_t *mx) {
    return guard([&] { RC_CHECK(mn && mx, RC_ERR_INVALID_VALUE, "null out"); bounds_index(from_c(l), mn, mx); });
}
int rc_layout_c_contig(const rc_layout *l, int *out) {
    return guard([&] { RC_CHECK(out, RC_ERR_INVALID_VALUE, "null out"); *out = c_contig(from_c(l)) ? 1 : 0; });
}
int rc_layout_f_contig(const rc_layout *l, int *out) {
    return guard([&] { RC_CHECK(out, RC_ERR_INVALID_VALUE, "null out"); *out = f_contig(from_c(l)) ? 1 : 0; });
}
int rc_layout_new_contig(const int64_t *shape, int ndim, rc_order order, int64_t offset, rc_layout *out) {
    return guard([&] {
        RC_CHECK(ndim >= 0 && ndim <= RC_MAX_NDIM && (ndim == 0 || shape), RC_ERR_INVALID_LAYOUT, "invalid shape");
        to_c(new_contig(std::vector<int64_t>(shape, shape + ndim), order, offset), out);
    });
}
int rc_layout_broadcast(const rc_layout *la, const rc_layout *lb, rc_order order, rc_layout *oa, rc_layout *ob) {
    return guard([&] {
        Layout ra, rb;
        broadcast_layouts(from_c(la), from_c(lb), order, &ra, &rb);
        to_c(ra, oa);
        to_c(rb, ob);
    });
}
int rc_layout_for_binary_op(const rc_layout *la, const rc_layout *lb, rc_order order, rc_layout *lc) {
    return guard([&] { to_c(layout_for_binary_op(from_c(la), from_c(lb), order), lc); });
}
int rc_layout_for_array_copy(const rc_layout *la, rc_iter_order it, rc_order order, rc_layout *lc) {
    return guard([&] { to_c(layout_for_array_copy(from_c(la), it, order), lc); });
}
int rc_layout_for_reduce(const rc_layout *la, const int64_t *axes, int naxes, rc_layout *lo) {
    return guard([&] {
        Layout l = from_c(la);
        to_c(layout_for_reduce(l, normalize_axes(axes, naxes, l.ndim())), lo);
    });
}
int rc_layout_reshapeable(const rc_layout *la, const int64_t *shape, int ndim, rc_order order, int *viewable,
                          rc_layout *out) {
    return guard([&] {
        RC_CHECK(viewable && out, RC_ERR_INVALID_VALUE, "null out");
        RC_CHECK(ndim >= 0 && ndim <= RC_MAX_NDIM && (ndim == 0 || shape), RC_ERR_INVALID_LAYOUT, "invalid shape");
        Layout r;
        bool ok = reshapeable(from_c(la), std::vector<int64_t>(shape, shape + ndim), order, &r);
        *viewable = ok ? 1 : 0;
        if (ok) to_c(r, out);
    });
}
int rc_layout_equal(const rc_layout *a, const rc_layout *b, int *equal) {
    return guard([&] { RC_CHECK(equal, RC_ERR_INVALID_VALUE, "null out"); *equal = layout_equal(from_c(a), from_c(b)) ? 1 : 0; });
}

/* ---------------- assign / fill ---------------- */
int rc_assign(rc_device *dev, rc_dtype tc, void *c, const rc_layout *lc_, rc_dtype ta, const void *a,
              const rc_layout *la_) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout lc = from_c(lc_), la = from_c(la_);
        CanonEw cn = canon_elementwise({&lc, &la}, false, true);
        if (cn.empty) return;
        check_ptr(c, "c"); check_ptr(a, "a");
        EwArgs args;
        args.c = c;
        args.a = a;
        run_cast(dev, tc, ta, cn, args);
    });
}

static int assign_arbitary_impl(rc_device *dev, const rc_order *order_in, rc_dtype tc, void *c, const rc_layout *lc_,
                               rc_dtype ta, const void *a, const rc_layout *la_) {
    return guard([&] {
        DeviceGuard g(dev);
        const rc_order order = order_in ? *order_in : dev->order;
        RC_CHECK(order == RC_ROW_MAJOR || order == RC_COL_MAJOR, RC_ERR_INVALID_VALUE, "invalid order");
        Layout lc = from_c(lc_), la = from_c(la_);
        RC_CHECK(lc.size() == la.size(), RC_ERR_INVALID_LAYOUT, "assign_arbitary requires layouts of equal size");
        if (lc.size() == 0) return;
        check_ptr(c, "c"); check_ptr(a, "a");
        Layout rc_, ra_;
        if (refine_to_common_shape(lc, la, order, &rc_, &ra_)) {
            CanonEw cn = canon_elementwise({&rc_, &ra_}, false);
            EwArgs args;
            args.c = c;
            args.a = a;
            run_cast(dev, tc, ta, cn, args);
        } else if (tc == ta) {
            run_assign_arbitrary_generic(dev, tc, c, lc, ta, a, la, order);
        } else {
            // rare: strided views of incompatible shapes AND a cast -- stage the cast through a flat buffer
            // (a 1-D contiguous layout has a common refinement with every shape)
            Layout flat = new_contig({lc.size()}, RC_ROW_MAJOR, 0);
            void *tmp = nullptr;
            cudaError_t e = cudaMallocAsync(&tmp, (size_t)lc.size() * dtype_size(tc), dev->stream);
            if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
            try {
                Layout f1, a1;
                RC_CHECK(refine_to_common_shape(flat, la, order, &f1, &a1), RC_ERR_RUNTIME, "refinement of a flat layout");
                CanonEw cn = canon_elementwise({&f1, &a1}, false);
                EwArgs args;
                args.c = tmp;
                args.a = a;
                run_cast(dev, tc, ta, cn, args);
                run_assign_arbitrary_generic(dev, tc, c, lc, tc, tmp, flat, order);
            } catch (...) {
                cudaFreeAsync(tmp, dev->stream);
                throw;
            }
            RC_CUDA(cudaFreeAsync(tmp, dev->stream));
        }
    });
}

int rc_assign_arbitary(rc_device *dev, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta, const void *a,
                       const rc_layout *la) {
    return assign_arbitary_impl(dev, nullptr, tc, c, lc, ta, a, la);
}
int rc_assign_arbitary_order(rc_device *dev, rc_order order, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta,
                             const void *a, const rc_layout *la) {
    return assign_arbitary_impl(dev, &order, tc, c, lc, ta, a, la);
}

int rc_fill(rc_device *dev, rc_dtype tc, void *c, const rc_layout *lc_, rc_dtype tf, const void *fill) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout lc = from_c(lc_);
        check_ptr(fill, "fill");
        CanonEw cn = canon_elementwise({&lc}, true);  // iteration order G (cpu_rayon/assignment.rs:199)
        if (cn.empty) return;
        check_ptr(c, "c");
        unsigned char v[16];
        host_scalar_cast(tc, tf, fill, v);
        run_fill(dev, tc, cn, c, v);
    });
}

/* ---------------- elementwise ---------------- */
int rc_op_mutc_refa_refb(rc_device *dev, rc_binop op, rc_dtype t, void *c, const rc_layout *lc_, const void *a,
                         const rc_layout *la_, const void *b, const rc_layout *lb_) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout lc = from_c(lc_), la = from_c(la_), lb = from_c(lb_);
        CanonEw cn = canon_elementwise({&lc, &la, &lb}, false, true);
        if (cn.empty) return;
        check_ptr(c, "c"); check_ptr(a, "a"); check_ptr(b, "b");
        EwArgs args;
        args.c = c; args.a = a; args.b = b;
        dispatch_binary(dev, op, t, cn, args);
    });
}

int rc_op_mutc_refa_numb(rc_device *dev, rc_binop op, rc_dtype t, void *c, const rc_layout *lc_, const void *a,
                         const rc_layout *la_, const void *b_host) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout lc = from_c(lc_), la = from_c(la_);
        CanonEw cn = canon_elementwise({&lc, &la}, false, true);
        if (cn.empty) return;
        check_ptr(c, "c"); check_ptr(a, "a"); check_ptr(b_host, "b");
        EwArgs args;
        args.c = c; args.a = a; args.b_const = true; args.b_host = b_host;
        dispatch_binary(dev, op, t, cn, args);
    });
}

int rc_op_mutc_numa_refb(rc_device *dev, rc_binop op, rc_dtype t, void *c, const rc_layout *lc_, const void *a_host,
                         const void *b, const rc_layout *lb_) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout lc = from_c(lc_), lb = from_c(lb_);
        CanonEw cn = canon_elementwise({&lc, &lb}, false, true);
        if (cn.empty) return;
        check_ptr(c, "c"); check_ptr(b, "b"); check_ptr(a_host, "a");
        EwArgs args;
        args.c = c; args.a_const = true; args.a_host = a_host; args.b = b;
        dispatch_binary(dev, op, t, cn, args);
    });
}

int rc_op_muta_refb(rc_device *dev, rc_binop op, rc_dtype t, void *a, const rc_layout *la_, const void *b,
                    const rc_layout *lb_, int reverse) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(!is_cmp(op), RC_ERR_INVALID_VALUE, "comparison ops have no in-place form");
        Layout la = from_c(la_), lb = from_c(lb_);
        check_ptr(a, "a"); check_ptr(b, "b");
        EwArgs args;
        args.c = a;
        CanonEw cn;
        if (!reverse) { cn = canon_elementwise({&la, &la, &lb}, false); args.a = a; args.b = b; }
        else          { cn = canon_elementwise({&la, &lb, &la}, false); args.a = b; args.b = a; }
        dispatch_binary(dev, op, t, cn, args);
    });
}

int rc_op_muta_numb(rc_device *dev, rc_binop op, rc_dtype t, void *a, const rc_layout *la_, const void *b_host,
                    int reverse) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(!is_cmp(op), RC_ERR_INVALID_VALUE, "comparison ops have no in-place form");
        Layout la = from_c(la_);
        check_ptr(b_host, "b");
        // iteration order G (cpu_rayon/op_with_func.rs:298): a broadcast axis of `a` is updated once
        CanonEw one = canon_elementwise({&la}, true);
        if (one.empty) return;
        check_ptr(a, "a");
        CanonEw cn = one;
        cn.nops = 2;
        cn.stride[1] = one.stride[0];
        cn.base[1] = one.base[0];
        EwArgs args;
        args.c = a;
        if (!reverse) { args.a = a; args.b_const = true; args.b_host = b_host; }
        else          { args.a_const = true; args.a_host = b_host; args.b = a; }
        dispatch_binary(dev, op, t, cn, args);
    });
}

int rc_unary_muta_refb(rc_device *dev, rc_unop op, rc_dtype t, void *a, const rc_layout *la_, const void *b,
                       const rc_layout *lb_) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout la = from_c(la_), lb = from_c(lb_);
        CanonEw cn = canon_elementwise({&la, &lb}, false, true);
        if (cn.empty) return;
        check_ptr(a, "a"); check_ptr(b, "b");
        EwArgs args;
        args.c = a; args.a = b;
        if (dtype_is_extended(t)) {
            RC_CHECK(run_unary_ext(dev, op, t, cn, args), RC_ERR_UNIMPLEMENTED,
                     std::string("this unary op is not implemented for dtype ") + dtype_name(t));
            return;
        }
        run_unary(dev, op, t, cn, args);
    });
}

int rc_unary_muta(rc_device *dev, rc_unop op, rc_dtype t, void *a, const rc_layout *la_) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(!is_predicate(op), RC_ERR_INVALID_VALUE, "boolean-output ops have no in-place form");
        Layout la = from_c(la_);
        CanonEw one = canon_elementwise({&la}, true);  // iteration order G (cpu_rayon/op_with_func.rs:355)
        if (one.empty) return;
        check_ptr(a, "a");
        CanonEw cn = one;
        cn.nops = 2;
        cn.stride[1] = one.stride[0];
        cn.base[1] = one.base[0];
        EwArgs args;
        args.c = a; args.a = a;
        if (dtype_is_extended(t)) {
            RC_CHECK(!(dtype_is_complex(t) && (op == RC_ABS || op == RC_REAL || op == RC_IMAG)), RC_ERR_INVALID_VALUE,
                     "abs / real / imag of a complex tensor change the element type: no in-place form");
            RC_CHECK(run_unary_ext(dev, op, t, cn, args), RC_ERR_UNIMPLEMENTED,
                     std::string("this unary op is not implemented for dtype ") + dtype_name(t));
            return;
        }
        run_unary(dev, op, t, cn, args);
    });
}

int rc_binop_out_dtype(rc_binop op, rc_dtype t, rc_dtype *out) {
    return guard([&] { RC_CHECK(out, RC_ERR_INVALID_VALUE, "null out"); *out = is_cmp(op) ? RC_BOOL : t; });
}
int rc_unop_out_dtype(rc_unop op, rc_dtype t, rc_dtype *out) {
    return guard([&] {
        RC_CHECK(out, RC_ERR_INVALID_VALUE, "null out");
        *out = is_predicate(op) ? RC_BOOL : t;
        // ExtNum::AbsOut of a complex number is its real type (abs / real / imag; auto_impl/op_binary_common.rs:122-204)
        if (dtype_is_complex(t) && (op == RC_ABS || op == RC_REAL || op == RC_IMAG)) *out = (t == RC_C32) ? RC_F32 : RC_F64;
    });
}
int rc_redop_out_dtype(rc_redop op, rc_dtype t, rc_dtype *out) {
    return guard([&] { RC_CHECK(out, RC_ERR_INVALID_VALUE, "null out"); *out = redop_out_dtype(op, t); });
}

/* ---------------- reductions ---------------- */
int rc_reduce_all_device(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const rc_layout *la_, void *dev_out) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout la = from_c(la_);
        check_ptr(dev_out, "dev_out");
        if (la.size() != 0) check_ptr(a, "a");
        Layout lo;  // 0-d
        reduce_into(dev, op, t, a, la, all_axes(la.ndim()), dev_out, lo);
    });
}

int rc_reduce_all(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const rc_layout *la_, void *host_out) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout la = from_c(la_);
        check_ptr(host_out, "host_out");
        if (la.size() != 0) check_ptr(a, "a");
        std::lock_guard<std::mutex> slot_lock(dev->slot_mu);
        void *host = nullptr;
        void *slot = scalar_slot(dev, &host);
        Layout lo;
        reduce_into(dev, op, t, a, la, all_axes(la.ndim()), slot, lo);
        RC_CUDA(cudaStreamSynchronize(dev->stream));
        std::memcpy(host_out, host, dtype_size(redop_out_dtype(op, t)));
    });
}

int rc_reduce_axes_into(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const rc_layout *la_,
                        const int64_t *axes, int naxes, void *out, const rc_layout *lo_) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout la = from_c(la_), lo = from_c(lo_);
        std::vector<int> ax = normalize_axes(axes, naxes, la.ndim());
        if (lo.size() != 0) check_ptr(out, "out");
        reduce_into(dev, op, t, a, la, ax, out, lo);
    });
}

int rc_reduce_axes(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const rc_layout *la_, const int64_t *axes,
                   int naxes, void **out_dev, rc_layout *lo_out) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(out_dev && lo_out, RC_ERR_INVALID_VALUE, "null out");
        Layout la = from_c(la_);
        std::vector<int> ax = normalize_axes(axes, naxes, la.ndim());
        if (op == RC_MAX || op == RC_MIN)
            RC_CHECK(la.size() != 0, RC_ERR_INVALID_VALUE,
                     op == RC_MAX ? "zero-size array is not supported for max" : "zero-size array is not supported for min");
        Layout lo = layout_for_reduce(la, ax);
        size_t nbytes = (size_t)std::max<int64_t>(lo.size(), 1) * dtype_size(redop_out_dtype(op, t));
        void *p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, nbytes, dev->stream);
        if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
        try {
            reduce_into(dev, op, t, a, la, ax, p, lo);
        } catch (...) {
            cudaFreeAsync(p, dev->stream);
            throw;
        }
        *out_dev = p;
        to_c(lo, lo_out);
    });
}

}  // extern "C"
