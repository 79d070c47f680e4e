// rc_create.cu -- device-side creation: arange, linspace, tril, triu.
//
// Replaces rstsr-native-impl/src/cpu_rayon/creation.rs:8-131 (arange by f64 / isize arithmetic, linspace) and
// cpu_serial/op_tri.rs:524-590 (tril / triu zero the other triangle of the last two axes in place).
// zeros / ones / full are rc_memset / rc_fill.
#include <cmath>

#include "rc_kernel_common.cuh"
#include "rc_layout.hpp"
#include "rc_types.cuh"

namespace rc {
namespace {

// arange_by_primitive_f64: T::from_f64(start + i as f64 * step) -- evaluated in f64 (no FMA contraction: -fmad=false)
template <class T>
__global__ void arange_f_kernel(T *out, int64_t n, double start, double step) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (T)(start + (double)i * step);
}
// arange_by_primitive_isize: T::from_isize(start + i * step)
template <class T>
__global__ void arange_i_kernel(T *out, int64_t n, int64_t start, int64_t step) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (T)(start + i * step);
}
// linspace: start + T::from(i) * step, in T (cpu_rayon/creation.rs:110-131)
template <class T>
__global__ void linspace_kernel(T *out, int64_t n, T start, T step) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = start + (T)i * step;
}

// the same for the half / complex element types, in THEIR arithmetic (half: every op through f32 and one rounding;
// complex: the textbook product of rc_types.cuh).  T::from(i): NumCast of the half types goes through f32, of Complex<R>
// to (R::from(i), 0).
template <class T> __device__ __forceinline__ T from_index(int64_t i) {
    if constexpr (is_half_t<T>::value) return T((float)i);
    else return T((typename real_of<T>::type)i, (typename real_of<T>::type)0);
}
template <class T>
__global__ void linspace_x_kernel(T *out, int64_t n, T start, T step) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = start + from_index<T>(i) * step;
}
template <class T> T host_from_count(int64_t v) {
    if constexpr (is_half_t<T>::value) return T((float)v);
    else return T((typename real_of<T>::type)v, (typename real_of<T>::type)0);
}
template <class T>
void linspace_x(rc_device *dev, void *p, const void *start, const void *end, int64_t n, int endpoint) {
    T s, e;
    std::memcpy(&s, start, sizeof(T));
    std::memcpy(&e, end, sizeof(T));
    // step = (end - start) / T::from(n - 1 | n) in T (cpu_rayon/creation.rs:121-124); n == 1 returns [start]
    const T step = (n == 1) ? host_from_count<T>(0) : (e - s) / host_from_count<T>(endpoint ? n - 1 : n);
    const unsigned B = 256, grid = (unsigned)((n + B - 1) / B);
    linspace_x_kernel<T><<<grid, B, 0, dev->stream>>>((T *)p, n, s, step);
}

// arange of the types that take the reference's generic fallback (arange_by_partial_ord_cpu_serial,
// cpu_serial/creation.rs:7-19): `while current < end { push(current); current = current + step }` in the element type --
// a sequential recurrence (half types round at every step), so it is evaluated on the host and uploaded.
constexpr int64_t kMaxSerialArange = 1 << 24;
template <class T, class Next>
std::vector<T> arange_serial(T start, T end, Next next) {
    std::vector<T> v;
    T cur = start;
    while (cur < end) {
        RC_CHECK((int64_t)v.size() < kMaxSerialArange, RC_ERR_INVALID_VALUE,
                 "arange of an 8- / 16-bit type does not terminate within 2^24 elements (step too small or wrapping)");
        v.push_back(cur);
        cur = next(cur);
    }
    return v;
}

struct alignas(16) W16 { uint64_t lo, hi; };  // a 16-byte element (c64) as raw words

struct TriDesc {
    int ndim;
    uint32_t total;
    FastDiv div[RC_MAX_NDIM];     // fastest-first: dim 0 = last axis (j), dim 1 = second-to-last (i), then batch
    int64_t stride[RC_MAX_NDIM];
    int64_t k;
    int lower;                    // 1: tril (zero j > i + k), 0: triu (zero j < i + k)
};

template <class U>
__global__ void __launch_bounds__(256) tri_kernel(const __grid_constant__ TriDesc d, U *a) {
    uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= d.total) return;
    int64_t off = 0, i = 0, j = 0;
    uint32_t t = idx;
    for (int k = 0; k < d.ndim; ++k) {
        uint32_t q, r;
        d.div[k].divmod(t, q, r);
        off += (int64_t)r * d.stride[k];
        if (k == 0) j = r;
        if (k == 1) i = r;
        t = q;
    }
    const bool zero = d.lower ? (j > i + d.k) : (j < i + d.k);
    if (zero) a[off] = U{};
}

template <class U>
void tri_launch(rc_device *dev, const TriDesc &d, void *a) {
    tri_kernel<U><<<(d.total + 255) / 256, 256, 0, dev->stream>>>(d, static_cast<U *>(a));
    after_launch(dev, "tri_kernel");
}

double host_as_f64(rc_dtype t, const void *p) {
    switch (t) {
        case RC_F64: { double v; std::memcpy(&v, p, 8); return v; }
        case RC_F32: { float v; std::memcpy(&v, p, 4); return v; }
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "not a float dtype");
}
int64_t host_as_i64(rc_dtype t, const void *p) {
    switch (t) {
        case RC_I32: { int32_t v; std::memcpy(&v, p, 4); return v; }
        case RC_I64: { int64_t v; std::memcpy(&v, p, 8); return v; }
        case RC_U32: { uint32_t v; std::memcpy(&v, p, 4); return v; }
        case RC_U64: { uint64_t v; std::memcpy(&v, p, 8); RC_CHECK(v <= (uint64_t)INT64_MAX, RC_ERR_INVALID_VALUE, "arange bound exceeds isize"); return (int64_t)v; }
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, "arange is not implemented for this dtype (complex numbers and bool have no ordering / step)");
}

void *dev_alloc(rc_device *dev, size_t nbytes) {
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, nbytes ? nbytes : 1, dev->stream);
    if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    return p;
}

}  // namespace
}  // namespace rc

using namespace rc;

extern "C" {

int rc_arange(rc_device *dev, rc_dtype t, const void *start, const void *end, const void *step, void **out_dev,
              int64_t *n_out) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(start && end && step && out_dev && n_out, RC_ERR_INVALID_VALUE, "null argument");
        int64_t n = 0;
        const unsigned B = 256;
        if (t == RC_I8 || t == RC_I16 || t == RC_U8 || t == RC_U16 || dtype_is_half(t)) {
            // generic fallback of the reference: a serial recurrence in the element type
            std::vector<unsigned char> bytes;
            auto pack = [&](const auto &v) {
                n = (int64_t)v.size();
                bytes.resize(v.size() * sizeof(v[0]));
                if (!v.empty()) std::memcpy(bytes.data(), v.data(), bytes.size());
            };
            auto ints = [&](auto zero) {
                using T = decltype(zero);
                T s, e, st;
                std::memcpy(&s, start, sizeof(T)); std::memcpy(&e, end, sizeof(T)); std::memcpy(&st, step, sizeof(T));
                pack(arange_serial<T>(s, e, [st](T c) { return (T)(c + st); }));  // wrapping, as release builds do
            };
            auto halves = [&](auto zero) {
                using T = decltype(zero);
                T s, e, st;
                std::memcpy(&s, start, sizeof(T)); std::memcpy(&e, end, sizeof(T)); std::memcpy(&st, step, sizeof(T));
                pack(arange_serial<T>(s, e, [st](T c) { return c + st; }));
            };
            switch (t) {
                case RC_I8: ints(int8_t()); break;
                case RC_I16: ints(int16_t()); break;
                case RC_U8: ints(uint8_t()); break;
                case RC_U16: ints(uint16_t()); break;
                case RC_F16: halves(h16()); break;
                default: halves(b16()); break;
            }
            void *p = dev_alloc(dev, bytes.size());
            if (!bytes.empty()) {
                RC_CUDA(cudaMemcpyAsync(p, bytes.data(), bytes.size(), cudaMemcpyHostToDevice, dev->stream));
                RC_CUDA(cudaStreamSynchronize(dev->stream));  // `bytes` is pageable and dies with this scope
            }
            *out_dev = p;
        } else if (dtype_is_float(t)) {
            const double s = host_as_f64(t, start), e = host_as_f64(t, end), st = host_as_f64(t, step);
            RC_CHECK(st != 0.0, RC_ERR_INVALID_VALUE, "arange step must not be zero");  // auto_impl/creation.rs:81
            double cnt = std::ceil((e - s) / st);
            n = (cnt > 0 && std::isfinite(cnt)) ? (int64_t)cnt : 0;
            // the interval is open on the right: drop a last element that rounding pushed onto / past `end`
            auto value = [&](int64_t i) { double v = s + (double)i * st; return t == RC_F32 ? (double)(float)v : v; };
            if (n > 0 && ((st > 0 && value(n - 1) >= e) || (st < 0 && value(n - 1) <= e))) --n;
            void *p = dev_alloc(dev, (size_t)n * dtype_size(t));
            if (n > 0) {
                unsigned grid = (unsigned)((n + B - 1) / B);
                if (t == RC_F64) arange_f_kernel<double><<<grid, B, 0, dev->stream>>>((double *)p, n, s, st);
                else arange_f_kernel<float><<<grid, B, 0, dev->stream>>>((float *)p, n, s, st);
                after_launch(dev, "arange_f_kernel");
            }
            *out_dev = p;
        } else {
            const int64_t s = host_as_i64(t, start), e = host_as_i64(t, end), st = host_as_i64(t, step);
            RC_CHECK(st != 0, RC_ERR_INVALID_VALUE, "arange step must not be zero");
            double cnt = std::ceil((double)(e - s) / (double)st);
            n = cnt > 0 ? (int64_t)cnt : 0;
            if (n > 0) {
                int64_t last = s + (n - 1) * st;
                if ((st > 0 && last >= e) || (st < 0 && last <= e)) --n;
            }
            void *p = dev_alloc(dev, (size_t)n * dtype_size(t));
            if (n > 0) {
                unsigned grid = (unsigned)((n + B - 1) / B);
                switch (t) {
                    case RC_I32: arange_i_kernel<int32_t><<<grid, B, 0, dev->stream>>>((int32_t *)p, n, s, st); break;
                    case RC_U32: arange_i_kernel<uint32_t><<<grid, B, 0, dev->stream>>>((uint32_t *)p, n, s, st); break;
                    case RC_I64: arange_i_kernel<int64_t><<<grid, B, 0, dev->stream>>>((int64_t *)p, n, s, st); break;
                    default: arange_i_kernel<uint64_t><<<grid, B, 0, dev->stream>>>((uint64_t *)p, n, s, st); break;
                }
                after_launch(dev, "arange_i_kernel");
            }
            *out_dev = p;
        }
        *n_out = n;
    });
}

int rc_linspace(rc_device *dev, rc_dtype t, const void *start, const void *end, int64_t n, int endpoint,
                void **out_dev) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(start && end && out_dev, RC_ERR_INVALID_VALUE, "null argument");
        RC_CHECK(n >= 0, RC_ERR_INVALID_VALUE, "negative length");
        RC_CHECK(dtype_is_float(t) || dtype_is_extended(t), RC_ERR_UNIMPLEMENTED,
                 "linspace requires a floating-point or complex dtype (T: ComplexFloat)");
        void *p = dev_alloc(dev, (size_t)n * dtype_size(t));
        *out_dev = p;
        if (n == 0) return;
        const unsigned B = 256, grid = (unsigned)((n + B - 1) / B);
        if (dtype_is_extended(t)) {
            switch (t) {
                case RC_F16: linspace_x<h16>(dev, p, start, end, n, endpoint); break;
                case RC_BF16: linspace_x<b16>(dev, p, start, end, n, endpoint); break;
                case RC_C32: linspace_x<c32>(dev, p, start, end, n, endpoint); break;
                default: linspace_x<c64>(dev, p, start, end, n, endpoint); break;
            }
        } else if (t == RC_F64) {
            double s, e;
            std::memcpy(&s, start, 8);
            std::memcpy(&e, end, 8);
            double step = (n == 1) ? 0.0 : (endpoint ? (e - s) / (double)(n - 1) : (e - s) / (double)n);
            linspace_kernel<double><<<grid, B, 0, dev->stream>>>((double *)p, n, s, step);
        } else {
            float s, e;
            std::memcpy(&s, start, 4);
            std::memcpy(&e, end, 4);
            float step = (n == 1) ? 0.0f : (endpoint ? (e - s) / (float)(n - 1) : (e - s) / (float)n);
            linspace_kernel<float><<<grid, B, 0, dev->stream>>>((float *)p, n, s, step);
        }
        after_launch(dev, "linspace_kernel");
    });
}

static int tri_impl(rc_device *dev, rc_dtype t, void *a, const rc_layout *l_, int64_t k, int lower) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout l = from_c(l_);
        RC_CHECK(l.ndim() >= 2, RC_ERR_INVALID_LAYOUT, "tril / triu need at least two axes");
        if (l.size() == 0) return;
        RC_CHECK(a != nullptr, RC_ERR_INVALID_VALUE, "null pointer: a");
        RC_CHECK(l.size() <= kMaxItemsPerLaunch, RC_ERR_UNIMPLEMENTED, "tril / triu are limited to 2^31 elements");
        TriDesc d;
        std::memset(&d, 0, sizeof(d));
        d.ndim = l.ndim();
        d.total = (uint32_t)l.size();
        d.k = k;
        d.lower = lower;
        for (int q = 0; q < l.ndim(); ++q) {
            int ax = l.ndim() - 1 - q;
            d.div[q] = FastDiv((uint32_t)l.shape[ax]);
            d.stride[q] = l.stride[ax];
        }
        char *p = static_cast<char *>(a) + l.offset * (int64_t)dtype_size(t);
        switch (dtype_size(t)) {
            case 1: tri_launch<uint8_t>(dev, d, p); break;
            case 2: tri_launch<uint16_t>(dev, d, p); break;
            case 4: tri_launch<uint32_t>(dev, d, p); break;
            case 8: tri_launch<uint64_t>(dev, d, p); break;
            default: tri_launch<W16>(dev, d, p); break;
        }
    });
}

int rc_tril(rc_device *dev, rc_dtype t, void *a, const rc_layout *l, int64_t k) { return tri_impl(dev, t, a, l, k, 1); }
int rc_triu(rc_device *dev, rc_dtype t, void *a, const rc_layout *l, int64_t k) { return tri_impl(dev, t, a, l, k, 0); }

}  // extern "C"
