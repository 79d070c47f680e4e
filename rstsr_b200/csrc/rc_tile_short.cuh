// rc_tile_short.cuh -- transposing COPY when one of the two tile axes is SHORT and that side of the copy is one flat
// contiguous range: `(n, k).T -> (k, n)` (de-interleave: the source is n*k consecutive elements, the output k long rows)
// and `(k, n).T -> (n, k)` (interleave: k long source rows, the output n*k consecutive elements); 1-, 2-, 4- and 8-byte
// elements, k up to 64.  These are the coordinate / per-shell table layouts (xyz <-> structure of arrays).
//
// Why a third tile kernel: ew_tile_rect_kernel handles these shapes for any functor and any strides, but pays an (x, y)
// decomposition per element and role (~40 instructions per element; ncu: issue-active 76 % at 43 % DRAM for f32,
// profiles/r01_ncu_rect_summary.json), which caps 4-byte elements at 4.1-4.4 TB/s and narrower ones far lower.  For a
// plain copy with a FLAT side nothing needs decomposing:
//   flat side   the tile is R*k consecutive elements: moved as 16-byte vectors by linear index between global memory and
//               a linear image in shared memory (no element type, no coordinates);
//   rows side   warp = one of the k rows, lane = VE consecutive long positions (one 16-byte vector in global memory),
//               VE element-sized shared accesses at stride k.  R is a power of two, so (row, chunk) is a shift and a mask.
// Bank conflicts of the stride-k accesses are broken by the layout, not by a pitch: the image is linear, with one pad
// word (two for 8-byte elements, which must stay 8-byte aligned) after every UNIT = the k * 16 bytes that belong to one
// lane of the rows side.  Lanes of the rows side are then 4 k + 1 words apart -- odd, so conflict-free for every k -- and
// every aligned 16-byte granule stays contiguous for the flat side (4 STS.32 / LDS.32; 2-way conflicts there for most k).
// The unit of a granule is one FastDiv per 16 bytes, (row, chunk) of a warp one FastDiv per 512 bytes: nothing per
// element.  About 12 instructions per 4-byte element instead of 40.
#pragma once
#include "rc_kernel_common.cuh"

namespace rc {

constexpr int SHORT_THREADS = 256;
constexpr int SHORT_WARPS = SHORT_THREADS / 32;
constexpr int SHORT_U = 5;
constexpr int SHORT_TILE_BYTES = SHORT_U * SHORT_THREADS * 16;  // payload of one tile: 20 KB
constexpr int SHORT_MAX_K = 64;

struct ShortDesc {
    uint32_t k, n;       // short extent (taken whole), long extent
    uint32_t R;          // long positions per tile: a multiple of 32 * (16 / S)
    uint32_t tiles;      // tiles along the long axis
    FastDiv div_tiles;
    FastDiv div_k;       // granule -> unit
    FastDiv div_ch;      // 32-lane chunks per row of the tile (R / (32 * VE)): warp job -> (row, chunk)
    int nbatch;
    FastDiv bdiv[KMAXD];
    int64_t bstride_flat[KMAXD], bstride_rows[KMAXD];  // elements
    int64_t srow;        // distance between consecutive rows on the rows side (elements)
    uint32_t total;      // tiles * batch
};

// all U global loads of a trip must be issued before the first shared-memory store.  Left alone, ptxas sinks each load
// next to its store (fewer live registers), which serialises the memory latencies (measured 6.3 -> 5.2 TB/s); it does not
// move memory operations across a warp barrier, and the barrier itself does not wait for the loads.
#define SHORT_LOADS_FIRST() __syncwarp()

template <int S> struct short_word;
template <> struct short_word<1> { using type = uint8_t; };
template <> struct short_word<2> { using type = uint16_t; };
template <> struct short_word<4> { using type = uint32_t; };
template <> struct short_word<8> { using type = uint64_t; };

// S = element bytes; VEC: 16-byte vectors on both sides (host checked the alignment), else element by element;
// DEINT: flat -> rows (the flat side is the source), else rows -> flat.
template <int S, bool VEC, bool DEINT>
__global__ void __launch_bounds__(SHORT_THREADS) ew_tile_short_kernel(const __grid_constant__ ShortDesc d,
                                                                      unsigned char *__restrict__ flat,
                                                                      unsigned char *__restrict__ rows) {
    using T = typename short_word<S>::type;
    constexpr int VE = VEC ? 16 / S : 1;
    constexpr int UV = 16 / S;                 // long positions per unit
    constexpr uint32_t PB = S == 8 ? 8 : 4;    // pad bytes per unit
    // independent global loads per thread and loop trip: SHORT_TILE_BYTES / (SHORT_THREADS * 16), so ONE trip moves a whole
    // tile -- a second, mostly empty trip costs a full memory latency per CTA (measured: 18 KB tiles with U = 4 lose 10 %)
    constexpr int U = SHORT_U;
    extern __shared__ __align__(16) unsigned char short_smem[];
    unsigned char *sm = short_smem;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint32_t t = blockIdx.x, tile;
    d.div_tiles.divmod(t, t, tile);
    int64_t off_flat = 0, off_rows = 0;
#pragma unroll
    for (int i = 0; i < KMAXD; ++i) {
        if (i >= d.nbatch) break;
        uint32_t q, r;
        d.bdiv[i].divmod(t, q, r);
        off_flat += (int64_t)r * d.bstride_flat[i];
        off_rows += (int64_t)r * d.bstride_rows[i];
        t = q;
    }
    const uint32_t x0 = tile * d.R;
    const uint32_t rem = min(d.R, d.n - x0);
    unsigned char *fp = flat + (off_flat + (int64_t)x0 * d.k) * S;
    unsigned char *rp = rows + (off_rows + x0) * S;
    const uint32_t kS = d.k * S;

    auto flat_phase = [&]() {
        if constexpr (VEC) {
            const uint32_t nvec = (rem * kS) >> 4;
            for (uint32_t v0 = tid; v0 - lane < nvec; v0 += SHORT_THREADS * U) {  // warp-uniform trip count
                Pack<uint32_t, 4> w[U];
                if constexpr (DEINT) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t v = v0 + SHORT_THREADS * u;
                        ld_stream_pred<uint32_t, 4>(w[u], reinterpret_cast<const uint32_t *>(fp + (size_t)v * 16), v < nvec);
                    }
                    SHORT_LOADS_FIRST();
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t v = v0 + SHORT_THREADS * u;
                        if (v < nvec) {
                            uint32_t *s = reinterpret_cast<uint32_t *>(sm + 16 * v + PB * d.div_k.div(v));
#pragma unroll
                            for (int c = 0; c < 4; ++c) s[c] = w[u].v[c];
                        }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t v = v0 + SHORT_THREADS * u;
                        if (v < nvec) {
                            const uint32_t *s = reinterpret_cast<const uint32_t *>(sm + 16 * v + PB * d.div_k.div(v));
#pragma unroll
                            for (int c = 0; c < 4; ++c) w[u].v[c] = s[c];
                            st_stream<uint32_t, 4>(reinterpret_cast<uint32_t *>(fp + (size_t)v * 16), w[u]);
                        }
                    }
                }
            }
        } else {
            const uint32_t ne = rem * d.k;
            for (uint32_t i0 = tid; i0 - lane < ne; i0 += SHORT_THREADS * U) {  // warp-uniform trip count
                Pack<T, 1> w[U];
                if constexpr (DEINT) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t i = i0 + SHORT_THREADS * u;
                        ld_stream_pred<T, 1>(w[u], reinterpret_cast<const T *>(fp) + i, i < ne);
                    }
                    SHORT_LOADS_FIRST();
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t i = i0 + SHORT_THREADS * u;
                        if (i < ne) *reinterpret_cast<T *>(sm + i * S + PB * d.div_k.div((i * S) >> 4)) = w[u].v[0];
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t i = i0 + SHORT_THREADS * u;
                        if (i < ne) {
                            w[u].v[0] = *reinterpret_cast<const T *>(sm + i * S + PB * d.div_k.div((i * S) >> 4));
                            st_stream<T, 1>(reinterpret_cast<T *>(fp) + i, w[u]);
                        }
                    }
                }
            }
        }
    };

    auto rows_phase = [&]() {
        const uint32_t nq = d.k * d.div_ch.d, UB = 16 * d.k + PB;  // warp jobs; bytes per padded unit
        // position x = (chunk * 32 + lane) * VE of the tile: unit x / UV, (x % UV) * k elements into it
        auto image = [&](uint32_t x, uint32_t y) { return (x / UV) * UB + ((x % UV) * d.k + y) * S; };
        // warp job q -> (row y, first position x of this lane); x = rem switches the lane off
        auto job = [&](uint32_t q, uint32_t &x, uint32_t &y) {
            uint32_t ch;
            d.div_ch.divmod(q, y, ch);
            x = q < nq ? ((ch << 5) + lane) * VE : rem;
        };
        for (uint32_t q0 = warp; q0 < nq; q0 += SHORT_WARPS * U) {
            Pack<T, VE> w[U];
            if constexpr (!DEINT) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t x, y;
                    job(q0 + SHORT_WARPS * u, x, y);
                    ld_stream_pred<T, VE>(w[u], reinterpret_cast<const T *>(rp + ((int64_t)y * d.srow + x) * S), x < rem);
                }
                SHORT_LOADS_FIRST();
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint32_t x, y;
                job(q0 + SHORT_WARPS * u, x, y);
                if (x < rem) {
                    uint32_t b = image(x, y);
                    if constexpr (DEINT) {
#pragma unroll
                        for (int c = 0; c < VE; ++c, b += kS) w[u].v[c] = *reinterpret_cast<const T *>(sm + b);
                        st_stream<T, VE>(reinterpret_cast<T *>(rp + ((int64_t)y * d.srow + x) * S), w[u]);
                    } else {
#pragma unroll
                        for (int c = 0; c < VE; ++c, b += kS) *reinterpret_cast<T *>(sm + b) = w[u].v[c];
                    }
                }
            }
        }
    };

    if constexpr (DEINT) {
        flat_phase();
        __syncthreads();
        rows_phase();
    } else {
        rows_phase();
        __syncthreads();
        flat_phase();
    }
}

}  // namespace rc
