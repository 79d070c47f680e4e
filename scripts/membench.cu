// membench.cu -- scratch microbenchmark (not product): which load/store width, cache hint, bytes in
// flight and grid shape reach the HBM roofline on B200 for copy (1R+1W), add (2R+1W) and sum (1R).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu && ./membench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct alignas(32) V32 { double v[4]; };
struct alignas(16) V16 { double v[2]; };

template <int HINT> __device__ __forceinline__ V32 ld32(const V32* p) {
    V32 r;
    if (HINT == 0) asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]),"=d"(r.v[1]),"=d"(r.v[2]),"=d"(r.v[3]) : "l"(p));
    else if (HINT == 1) asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]),"=d"(r.v[1]),"=d"(r.v[2]),"=d"(r.v[3]) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]),"=d"(r.v[1]),"=d"(r.v[2]),"=d"(r.v[3]) : "l"(p));
    return r;
}
template <int HINT> __device__ __forceinline__ void st32(V32* p, const V32& r) {
    if (HINT == 1) asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p),"d"(r.v[0]),"d"(r.v[1]),"d"(r.v[2]),"d"(r.v[3]));
    else asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p),"d"(r.v[0]),"d"(r.v[1]),"d"(r.v[2]),"d"(r.v[3]));
}
template <int HINT> __device__ __forceinline__ V16 ld16(const V16* p) {
    V16 r;
    if (HINT == 0) asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]),"=d"(r.v[1]) : "l"(p));
    else if (HINT == 1) asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]),"=d"(r.v[1]) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]),"=d"(r.v[1]) : "l"(p));
    return r;
}
template <int HINT> __device__ __forceinline__ void st16(V16* p, const V16& r) {
    if (HINT == 1) asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p),"d"(r.v[0]),"d"(r.v[1]));
    else asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(p),"d"(r.v[0]),"d"(r.v[1]));
}

// NIN = 1 copy, 2 add.  W = 16 or 32 bytes per access.  U accesses per thread per array.  PERSIST: grid-stride.
template <int NIN, int W, int U, int HINT, bool PERSIST>
__global__ void __launch_bounds__(256) k_ew(const double* a, const double* b, double* c, size_t n_vec) {
    size_t tile = (size_t)blockDim.x * U;
    size_t ntiles = (n_vec + tile - 1) / tile;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        size_t base = t * tile + threadIdx.x;
        if (W == 32) {
            V32 x[U], y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < n_vec) { x[u] = ld32<HINT>((const V32*)a + i); if (NIN == 2) y[u] = ld32<HINT>((const V32*)b + i); } }
#pragma unroll
            for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < n_vec) { if (NIN == 2) { for (int j = 0; j < 4; ++j) x[u].v[j] += y[u].v[j]; } st32<HINT>((V32*)c + i, x[u]); } }
        } else {
            V16 x[U], y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < n_vec) { x[u] = ld16<HINT>((const V16*)a + i); if (NIN == 2) y[u] = ld16<HINT>((const V16*)b + i); } }
#pragma unroll
            for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < n_vec) { if (NIN == 2) { for (int j = 0; j < 2; ++j) x[u].v[j] += y[u].v[j]; } st16<HINT>((V16*)c + i, x[u]); } }
        }
        if (!PERSIST) break;
    }
}

template <int W, int U, int HINT>
__global__ void __launch_bounds__(256) k_sum(const double* a, double* out, size_t n_vec) {
    double acc[U * (W / 8)];
    for (int i = 0; i < U * (W / 8); ++i) acc[i] = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n_vec; i += U * stride) {
        if (W == 32) {
            V32 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = ld32<HINT>((const V32*)a + i + u * stride);
#pragma unroll
            for (int u = 0; u < U; ++u) for (int j = 0; j < 4; ++j) acc[u * 4 + j] += x[u].v[j];
        } else {
            V16 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = ld16<HINT>((const V16*)a + i + u * stride);
#pragma unroll
            for (int u = 0; u < U; ++u) for (int j = 0; j < 2; ++j) acc[u * 2 + j] += x[u].v[j];
        }
    }
    double s = 0;
    for (int k = 0; k < U * (W / 8); ++k) s += acc[k];
    for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(~0u, s, m);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

template <class F> float timeit(F f, int iters = 10) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / iters * 1e-3f;
}

int main() {
    const size_t N = 8192ull * 8192ull;  // f64 elements per array (512 MiB)
    double *a, *b, *c, *o;
    CK(cudaMalloc(&a, N * 8)); CK(cudaMalloc(&b, N * 8)); CK(cudaMalloc(&c, N * 8)); CK(cudaMalloc(&o, 8));
    CK(cudaMemset(a, 0, N * 8)); CK(cudaMemset(b, 0, N * 8));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    { float s = timeit([&] { cudaMemcpyAsync(c, a, N * 8, cudaMemcpyDeviceToDevice); }); printf("cudaMemcpy D2D          : %8.1f GB/s\n", 2.0 * N * 8 / s / 1e9); }

#define RUN_EW(NIN, W, U, HINT, PERSIST, GRIDMUL)                                                          \
    {                                                                                                      \
        size_t n_vec = N * 8 / W;                                                                          \
        size_t tiles = (n_vec + 256 * U - 1) / (256 * U);                                                  \
        unsigned grid = PERSIST ? (unsigned)(sms * GRIDMUL) : (unsigned)tiles;                             \
        float s = timeit([&] { k_ew<NIN, W, U, HINT, PERSIST><<<grid, 256>>>(a, b, c, n_vec); });         \
        printf("%s W=%2d U=%d hint=%d %s grid=%-7u: %8.1f GB/s\n", NIN == 1 ? "copy" : "add ", W, U, HINT, \
               PERSIST ? "persist" : "oneshot", grid, (NIN + 1.0) * N * 8 / s / 1e9);                      \
    }
    RUN_EW(1, 16, 1, 0, false, 0) RUN_EW(1, 16, 2, 0, false, 0) RUN_EW(1, 16, 4, 0, false, 0) RUN_EW(1, 16, 8, 0, false, 0)
    RUN_EW(1, 32, 1, 0, false, 0) RUN_EW(1, 32, 2, 0, false, 0) RUN_EW(1, 32, 4, 0, false, 0)
    RUN_EW(1, 16, 4, 1, false, 0) RUN_EW(1, 32, 2, 1, false, 0) RUN_EW(1, 32, 4, 1, false, 0)
    RUN_EW(1, 16, 4, 2, false, 0) RUN_EW(1, 32, 2, 2, false, 0) RUN_EW(1, 32, 4, 2, false, 0)
    RUN_EW(1, 16, 4, 0, true, 4) RUN_EW(1, 16, 4, 0, true, 8) RUN_EW(1, 32, 2, 0, true, 8) RUN_EW(1, 32, 4, 0, true, 4) RUN_EW(1, 32, 4, 0, true, 8)
    RUN_EW(1, 32, 2, 2, true, 8) RUN_EW(1, 32, 4, 2, true, 8)
    RUN_EW(2, 16, 1, 0, false, 0) RUN_EW(2, 16, 2, 0, false, 0) RUN_EW(2, 16, 4, 0, false, 0)
    RUN_EW(2, 32, 1, 0, false, 0) RUN_EW(2, 32, 2, 0, false, 0) RUN_EW(2, 32, 4, 0, false, 0)
    RUN_EW(2, 16, 4, 1, false, 0) RUN_EW(2, 32, 2, 1, false, 0)
    RUN_EW(2, 16, 4, 2, false, 0) RUN_EW(2, 32, 2, 2, false, 0) RUN_EW(2, 32, 4, 2, false, 0)
    RUN_EW(2, 16, 4, 0, true, 8) RUN_EW(2, 32, 2, 0, true, 8) RUN_EW(2, 32, 2, 2, true, 8) RUN_EW(2, 32, 2, 2, true, 4)

#define RUN_SUM(W, U, HINT, GRIDMUL)                                                                       \
    {                                                                                                      \
        size_t n_vec = N * 8 / W;                                                                          \
        unsigned grid = (unsigned)(sms * GRIDMUL);                                                         \
        float s = timeit([&] { k_sum<W, U, HINT><<<grid, 256>>>(a, o, n_vec); });                          \
        printf("sum  W=%2d U=%d hint=%d grid=%-7u        : %8.1f GB/s\n", W, U, HINT, grid, 1.0 * N * 8 / s / 1e9); \
    }
    RUN_SUM(16, 4, 0, 8) RUN_SUM(16, 8, 0, 8) RUN_SUM(32, 2, 0, 8) RUN_SUM(32, 4, 0, 8) RUN_SUM(32, 4, 0, 4)
    RUN_SUM(16, 4, 2, 8) RUN_SUM(32, 4, 2, 8) RUN_SUM(16, 4, 1, 8) RUN_SUM(32, 4, 1, 8) RUN_SUM(32, 4, 2, 16) RUN_SUM(32, 8, 2, 8)
    return 0;
}
