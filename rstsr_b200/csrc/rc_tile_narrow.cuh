// rc_tile_narrow.cuh -- permuted COPY of 1- and 2-byte elements.
//
// The generic tile kernel moves ONE element per lane and instruction, so a byte transpose issues eight times the
// instructions of an f64 transpose for the same bytes (measured round 1: u8 1.4-2.2 TB/s, i16 4.3 TB/s).  Here every
// global access is a 32-bit word of E = 4 / itemsize elements, on both sides:
//   phase 1  lanes run along Y (the source's unit-stride axis) in WORDS: a warp reads 2 x 128 contiguous bytes of one
//            source row per instruction pair and parks the words in shared memory at s[x][w ^ ((x / E) & 31)]
//            (32 + 32 words per row; the XOR spreads the rows a lane group reads later over all banks);
//   phase 2  a lane owns E consecutive x (one output word) and E consecutive y (one source word column): it reads
//            the E words s[E l + i][w], transposes the E x E element block in registers (PRMT) and stores E output
//            words, one per output row; lanes l = 0..31 cover 128 contiguous output bytes per row.
// Tile = (32 E) x (64 E) elements: 128 output bytes x 256 source bytes per row, 32 KiB (bytes) / 16 KiB (shorts).
// Eligibility (host side): same-size copy with one staged operand, extents and every stride that matters a multiple
// of E elements, 4-byte aligned bases.  Partial tiles are predicated per word.
#pragma once
#include "rc_kernel_common.cuh"

namespace rc {

constexpr int NW_WARPS = 8;
constexpr int NW_WORDS_Y = 64;   // source words per tile row
constexpr int NW_WORDS_X = 32;   // output words per tile row

struct NarrowDesc {
    uint32_t nx, ny;             // extents in ELEMENTS
    uint32_t tiles_x, tiles_y;
    int nbatch;
    uint32_t total_tiles;
    FastDiv div_ty, div_tx;
    FastDiv bdiv[KMAXD];
    int64_t bstride_c[KMAXD], bstride_a[KMAXD];  // elements
    int64_t sx_a, sy_c;                           // source stride along X, output stride along Y (elements)
};

// E x E element transpose of E words (row i = word w[i], element j of a word = bits [j * 8 * ESZ, ...))
template <int ESZ>
__device__ __forceinline__ void nw_transpose(uint32_t (&w)[4 / ESZ]) {
    if constexpr (ESZ == 2) {
        const uint32_t a = w[0], b = w[1];
        w[0] = __byte_perm(a, b, 0x5410);  // {a.lo, b.lo}
        w[1] = __byte_perm(a, b, 0x7632);  // {a.hi, b.hi}
    } else {
        const uint32_t t0 = __byte_perm(w[0], w[1], 0x5140);  // a0 b0 a1 b1
        const uint32_t t1 = __byte_perm(w[0], w[1], 0x7362);  // a2 b2 a3 b3
        const uint32_t t2 = __byte_perm(w[2], w[3], 0x5140);  // c0 d0 c1 d1
        const uint32_t t3 = __byte_perm(w[2], w[3], 0x7362);  // c2 d2 c3 d3
        w[0] = __byte_perm(t0, t2, 0x5410);                   // a0 b0 c0 d0
        w[1] = __byte_perm(t0, t2, 0x7632);                   // a1 b1 c1 d1
        w[2] = __byte_perm(t1, t3, 0x5410);                   // a2 b2 c2 d2
        w[3] = __byte_perm(t1, t3, 0x7632);                   // a3 b3 c3 d3
    }
}

template <int ESZ>
__global__ void __launch_bounds__(NW_WARPS * 32) ew_tile_narrow_kernel(const __grid_constant__ NarrowDesc d,
                                                                       unsigned char *__restrict__ c,
                                                                       const unsigned char *__restrict__ a) {
    constexpr int E = 4 / ESZ;                  // elements per word
    constexpr int TX = NW_WORDS_X * E;          // tile extent along X (elements)
    constexpr int TY = NW_WORDS_Y * E;          // tile extent along Y (elements)
    constexpr int ROWS_PER_WARP = TX / NW_WARPS;
    __shared__ uint32_t s[TX * NW_WORDS_Y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t t = blockIdx.x, ty, tx;
    d.div_ty.divmod(t, t, ty);
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
    const uint32_t remx = d.nx - x0, remy = d.ny - y0;       // valid elements of this tile
    const uint32_t wy_valid = remy / E < (uint32_t)NW_WORDS_Y ? remy / E : NW_WORDS_Y;  // ny % E == 0
    const uint32_t wx_valid = remx / E < (uint32_t)NW_WORDS_X ? remx / E : NW_WORDS_X;  // nx % E == 0

    // phase 1: source rows x = warp, warp + 8, ...; two words per lane and row
    const unsigned char *pa = a + (base_a + (int64_t)x0 * d.sx_a + y0) * ESZ;
    Pack<uint32_t, 1> r0[ROWS_PER_WARP], r1[ROWS_PER_WARP];
#pragma unroll
    for (int k = 0; k < ROWS_PER_WARP; ++k) {
        const uint32_t x = warp + k * NW_WARPS;
        const uint32_t *row = reinterpret_cast<const uint32_t *>(pa + (int64_t)x * d.sx_a * ESZ);
        r0[k].v[0] = 0;
        r1[k].v[0] = 0;
        ld_stream_pred<uint32_t, 1>(r0[k], row + lane, x < remx && (uint32_t)lane < wy_valid);
        ld_stream_pred<uint32_t, 1>(r1[k], row + lane + 32, x < remx && (uint32_t)(lane + 32) < wy_valid);
    }
#pragma unroll
    for (int k = 0; k < ROWS_PER_WARP; ++k) {
        const uint32_t x = warp + k * NW_WARPS;
        const uint32_t sw = (x / E) & 31u;
        s[x * NW_WORDS_Y + ((uint32_t)lane ^ sw)] = r0[k].v[0];
        s[x * NW_WORDS_Y + 32 + ((uint32_t)lane ^ sw)] = r1[k].v[0];
    }
    __syncthreads();

    // phase 2: word columns w = warp, warp + 8, ... of the source tile; lane l owns x = E l .. E l + E - 1
    unsigned char *pc = c + (base_c + (int64_t)y0 * d.sy_c + x0) * ESZ;
    const uint32_t sw = (uint32_t)lane & 31u;  // ((E * lane + i) / E) & 31
#pragma unroll
    for (int k = 0; k < NW_WORDS_Y / NW_WARPS; ++k) {
        const uint32_t w = warp + k * NW_WARPS;   // source word column: y = E w .. E w + E - 1
        uint32_t v[E];
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = s[(E * lane + i) * NW_WORDS_Y + (w & 32u) + ((w & 31u) ^ sw)];
        nw_transpose<ESZ>(v);
        if (w < wy_valid && (uint32_t)lane < wx_valid) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
                uint32_t *dst = reinterpret_cast<uint32_t *>(pc + (int64_t)(E * w + j) * d.sy_c * ESZ) + lane;
                __stcs(dst, v[j]);
            }
        }
    }
}

}  // namespace rc
