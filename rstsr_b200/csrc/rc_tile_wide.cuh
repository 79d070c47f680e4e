// rc_tile_wide.cuh -- permuted COPY of 8-byte elements through a WIDE shared-memory tile.
//
// Why: scripts/probe_runlen.py (profiles/r02_results/probe_runlen.json) measures what a contiguous run of a given length
// is worth on B200 for plain strided copies: 512-byte runs cost 2.8 % on the read side and 4.9 % on the write side,
// 1 KiB runs 1.2 % / 0.3 %, 256-byte runs 5.1 % / 6.6 % (against 4 KiB runs).  The square 64 x 64 tile moves 512-byte
// runs on both sides: 0.972 x 0.951 = 0.924 of the contiguous copy -- exactly the 6.42-6.48 TB/s it measures.  A tile
// that is wider along X (the output's contiguous axis) buys longer WRITE runs, the costlier side, at the same
// shared-memory footprint.
//
// Measured (scripts/probe_tile_bulk.py, profiles/r02_tile_wide.md): 128 x 32 with 256 threads (8 elements per thread) is
// +1.2 % on cfg2, +2.9 % on a 16384^2 transpose, +3.1 % on (512,512,2048) perm (1,2,0); 128 x 64 / 256 x 32 with 512
// threads are within 0.5 % of it, any shape with 16 or 32 elements per thread is 15 % slower.
//
// TX x TY tile, NT threads, same discipline as ew_tile_kernel: every load of the thread (predicated, streaming) is
// issued before the first shared-memory store; pitch TY + 1 keeps the column reads of phase 2 conflict-free per
// half-warp.  Copies only (one staged operand, no functor): everything else stays on ew_tile_kernel.
#pragma once
#include "rc_kernel_common.cuh"

namespace rc {

struct WideDesc {
    uint32_t nx, ny;
    uint32_t tiles_x, tiles_y;
    int nbatch;
    uint32_t total_tiles;
    FastDiv div_ty, div_tx;
    FastDiv bdiv[KMAXD];
    int64_t bstride_c[KMAXD], bstride_a[KMAXD];
    int64_t sx_a, sy_c;  // source stride along X, output stride along Y (elements); source sy = output sx = 1
};

template <class T, int TX, int TY, int NT>
__global__ void __launch_bounds__(NT) ew_tile_wide_kernel(const __grid_constant__ WideDesc d, T *__restrict__ c,
                                                           const T *__restrict__ a) {
    constexpr int NW = NT / 32;
    static_assert(TX % 32 == 0 && TY % 32 == 0 && TX % NW == 0 && TY % NW == 0, "tile shape");
    constexpr int PITCH = TY + 1;
    constexpr int R1 = TX / NW, C1 = TY / 32;  // phase 1: rows x per warp, 32-lane columns along y
    constexpr int R2 = TY / NW, C2 = TX / 32;  // phase 2: rows y per warp, 32-lane columns along x
    extern __shared__ __align__(16) unsigned char wide_smem[];
    T *s = reinterpret_cast<T *>(wide_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t t = blockIdx.x, ty, tx;
    d.div_ty.divmod(t, t, ty);  // consecutive blocks step along Y: adjacent READ runs
    d.div_tx.divmod(t, t, tx);
    int64_t base_c = 0, base_a = 0;
#pragma unroll
    for (int i = 0; i < KMAXD; ++i) {
        if (i >= d.nbatch) break;
        uint32_t q, r;
        d.bdiv[i].divmod(t, q, r);
        base_c += (int64_t)r * d.bstride_c[i];
        base_a += (int64_t)r * d.bstride_a[i];
        t = q;
    }
    const uint32_t x0 = tx * TX, y0 = ty * TY;
    const uint32_t remx = d.nx - x0, remy = d.ny - y0;

    const T *pa = a + base_a + (int64_t)x0 * d.sx_a + y0;
    Pack<T, 1> v[R1][C1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
        const uint32_t x = warp + NW * r;
#pragma unroll
        for (int j = 0; j < C1; ++j) {
            const uint32_t y = lane + 32 * j;
            v[r][j].v[0] = T();
            ld_stream_pred<T, 1>(v[r][j], pa + (int64_t)x * d.sx_a + y, x < remx && y < remy);
        }
    }
    __syncwarp();  // scheduling fence: ptxas keeps every load above it (see SHORT_LOADS_FIRST in rc_tile_short.cuh)
#pragma unroll
    for (int r = 0; r < R1; ++r)
#pragma unroll
        for (int j = 0; j < C1; ++j) s[(warp + NW * r) * PITCH + lane + 32 * j] = v[r][j].v[0];
    __syncthreads();

    T *pc = c + base_c + (int64_t)y0 * d.sy_c + x0;
#pragma unroll
    for (int r = 0; r < R2; ++r) {
        const uint32_t y = warp + NW * r;
        T *row = pc + (int64_t)y * d.sy_c;
#pragma unroll
        for (int j = 0; j < C2; ++j) {
            const uint32_t x = lane + 32 * j;
            if (y < remy && x < remx) {
                Pack<T, 1> o;
                o.v[0] = s[x * PITCH + y];
                st_stream<T, 1>(row + x, o);
            }
        }
    }
}

}  // namespace rc
