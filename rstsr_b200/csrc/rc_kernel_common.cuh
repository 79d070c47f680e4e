// rc_kernel_common.cuh -- device-side helpers shared by the kernel families (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rc_canon.hpp"
#include "rc_device.hpp"

namespace rc {

// ---- division by a runtime-constant 32-bit divisor (numerators < 2^31) ----
// q = (umulhi(n, m) + n) >> s with s = ceil(log2 d), m = floor(2^32 (2^s - d) / d) + 1.
struct FastDiv {
    uint32_t d, m, s;
    FastDiv() : d(1), m(1), s(0) {}
    explicit FastDiv(uint32_t div) : d(div) {
        s = 0;
        while ((1ull << s) < d) ++s;
        m = (uint32_t)(((1ull << 32) * ((1ull << s) - d)) / d + 1);
    }
    __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
        return (__umulhi(n, m) + n) >> s;
#else
        return (uint32_t)((((uint64_t)n * m) >> 32) + n) >> s;
#endif
    }
    __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &r) const {
        q = div(n);
        r = n - q * d;
    }
};

// ---- 16-byte-chunked vector of N elements of T (N * sizeof(T) is 1,2,4,8,16 or a multiple of 16) ----
template <typename T, int N>
struct alignas((N * sizeof(T)) >= 32 ? 32 : ((N * sizeof(T)) >= 16 ? 16 : (N * sizeof(T)))) Pack {
    T v[N];
};

// streaming (read-once) global load / store of a whole Pack with the widest instructions available:
// 256-bit LDG/STG (sm_100a) for 32-byte packs, 128-bit below.  The pointer must be aligned to the pack size.
template <typename T, int N>
__device__ __forceinline__ Pack<T, N> ld_stream(const T *p) {
    constexpr int BYTES = N * sizeof(T);
    Pack<T, N> r;
    if constexpr (BYTES == 32) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(&r);
        asm volatile("ld.global.cs.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(o[0]), "=l"(o[1]), "=l"(o[2]), "=l"(o[3]) : "l"(p));
    } else if constexpr (BYTES >= 16) {
        static_assert(BYTES % 16 == 0, "pack size");
        const int4 *q = reinterpret_cast<const int4 *>(p);
        int4 *o = reinterpret_cast<int4 *>(&r);
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) o[i] = __ldcs(q + i);
    } else if constexpr (BYTES == 8) {
        *reinterpret_cast<int2 *>(&r) = __ldcs(reinterpret_cast<const int2 *>(p));
    } else if constexpr (BYTES == 4) {
        *reinterpret_cast<int *>(&r) = __ldcs(reinterpret_cast<const int *>(p));
    } else if constexpr (BYTES == 2) {
        *reinterpret_cast<short *>(&r) = __ldcs(reinterpret_cast<const short *>(p));
    } else {
        *reinterpret_cast<char *>(&r) = __ldcs(reinterpret_cast<const char *>(p));
    }
    return r;
}

template <typename T, int N>
__device__ __forceinline__ void st_stream(T *p, const Pack<T, N> &r) {
    constexpr int BYTES = N * sizeof(T);
    if constexpr (BYTES == 32) {
        const unsigned long long *o = reinterpret_cast<const unsigned long long *>(&r);
        asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(o[0]), "l"(o[1]), "l"(o[2]), "l"(o[3])
                     : "memory");
    } else if constexpr (BYTES >= 16) {
        int4 *q = reinterpret_cast<int4 *>(p);
        const int4 *o = reinterpret_cast<const int4 *>(&r);
#pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) __stcs(q + i, o[i]);
    } else if constexpr (BYTES == 8) {
        __stcs(reinterpret_cast<int2 *>(p), *reinterpret_cast<const int2 *>(&r));
    } else if constexpr (BYTES == 4) {
        __stcs(reinterpret_cast<int *>(p), *reinterpret_cast<const int *>(&r));
    } else if constexpr (BYTES == 2) {
        __stcs(reinterpret_cast<short *>(p), *reinterpret_cast<const short *>(&r));
    } else {
        __stcs(reinterpret_cast<char *>(p), *reinterpret_cast<const char *>(&r));
    }
}

// Predicated in-place streaming load: `r` keeps its previous contents when `pred` is false.  Written as one
// predicated LDG so that (a) no branch surrounds the load and (b) the compiler cannot fold a later select into
// a MOV that waits on the load -- both serialise a thread's loads (seen in SASS; cost 25 % of bandwidth).
template <typename T, int N>
__device__ __forceinline__ void ld_stream_pred(Pack<T, N> &r, const T *p, bool pred) {
    constexpr int BYTES = N * sizeof(T);
    const int pi = pred ? 1 : 0;
    if constexpr (BYTES == 32) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(&r);
        asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %5, 0;\n @q ld.global.cs.v4.b64 {%0,%1,%2,%3}, [%4];\n}"
                     : "+l"(o[0]), "+l"(o[1]), "+l"(o[2]), "+l"(o[3]) : "l"(p), "r"(pi));
    } else if constexpr (BYTES == 16) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(&r);
        asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n @q ld.global.cs.v2.b64 {%0,%1}, [%2];\n}"
                     : "+l"(o[0]), "+l"(o[1]) : "l"(p), "r"(pi));
    } else if constexpr (BYTES == 8) {
        unsigned long long *o = reinterpret_cast<unsigned long long *>(&r);
        asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.cs.b64 %0, [%1];\n}"
                     : "+l"(o[0]) : "l"(p), "r"(pi));
    } else if constexpr (BYTES == 4) {
        unsigned int *o = reinterpret_cast<unsigned int *>(&r);
        asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.cs.b32 %0, [%1];\n}"
                     : "+r"(o[0]) : "l"(p), "r"(pi));
    } else if constexpr (BYTES == 2) {
        unsigned short *o = reinterpret_cast<unsigned short *>(&r);
        asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.cs.b16 %0, [%1];\n}"
                     : "+h"(o[0]) : "l"(p), "r"(pi));
    } else {
        static_assert(BYTES == 1, "pack size");
        unsigned short t = *reinterpret_cast<unsigned char *>(&r);
        asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.cs.u8 %0, [%1];\n}"
                     : "+h"(t) : "l"(p), "r"(pi));
        *reinterpret_cast<unsigned char *>(&r) = (unsigned char)t;
    }
}

// ---- per-launch descriptor of an elementwise kernel: NOPS operands over <= KMAXD merged dims ----
template <int NOPS>
struct EwDesc {
    int ndim;
    uint32_t total;           // work items (dim 0 counted in packs)
    FastDiv div[KMAXD];       // extents (dim 0 in packs)
    int64_t stride[NOPS][KMAXD];  // elements per unit step of each dim (dim 0: per PACK)
};

template <int NOPS>
__device__ __forceinline__ void ew_offsets(const EwDesc<NOPS> &d, uint32_t idx, int64_t (&off)[NOPS]) {
#pragma unroll
    for (int k = 0; k < NOPS; ++k) off[k] = 0;
#pragma unroll
    for (int i = 0; i < KMAXD; ++i) {
        if (i >= d.ndim) break;
        uint32_t q, r;
        if (i + 1 < d.ndim) {
            d.div[i].divmod(idx, q, r);
        } else {
            q = 0;
            r = idx;
        }
#pragma unroll
        for (int k = 0; k < NOPS; ++k) off[k] += (int64_t)r * d.stride[k][i];
        idx = q;
    }
}

constexpr int64_t kMaxItemsPerLaunch = (1ll << 31) - (1 << 20);  // keeps every 32-bit index < 2^31

}  // namespace rc
