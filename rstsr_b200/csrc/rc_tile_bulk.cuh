// rc_tile_bulk.cuh -- permuted COPY through TMA (cp.async.bulk.tensor, SASS UTMALDG / UTMASTG): the alternative staging
// the north star names next to the LDG / STG shared-memory tile of rc_elementwise.cuh ("or through TMA where the tensor
// is <= 5-D").
//
// One persistent CTA per SM walks 64 x 64 tiles of 8-byte elements:
//   load   ONE tensor-map box load per tile: 64 rows of the source (512 contiguous bytes each, along Y = the source's
//          unit-stride axis) land densely in shared memory and signal an mbarrier (expect_tx = 32 KiB); BK_STAGES tiles
//          are in flight per SM at any time, independent of what the threads do;
//   turn   the 256 threads transpose the tile shared -> shared: two 8-byte LDS along Y (conflict-free on the dense
//          tile), one 16-byte STS into the output tile, which is laid out as TMA's 128-byte swizzle expects (four
//          sub-tiles of 16 x-elements = 128-byte rows; 16-byte chunk c of row y sits at chunk c ^ (y & 7)), so the eight
//          lanes of a quarter-warp -- consecutive y, same x pair -- hit eight different bank groups;
//   store  FOUR tensor-map box stores per tile (16 x 64 elements each, SWIZZLE_128B) in one bulk group; the buffer is
//          reused once cp.async.bulk.wait_group.read says the engine has read it.
// No LDG / STG is issued for the payload.
//
// A first version moved the rows with 1-D bulk copies (cp.async.bulk, UBLKCP: 64 x 512-byte copies per tile and
// direction).  It was bit-exact and ran at 2.3 TB/s: the engine's per-copy cost (~40 cycles) caps 512-byte copies at
// ~13 bytes per cycle and SM (profiles/r02_tile_bulk.md).  Box copies are one descriptor per 8-32 KiB.
//
// Eligibility (host side, ew_launch_part + tile_tma_prepare): same-size 8-byte copy with one staged operand, nx and ny
// multiples of 64, at most three batch dims, every stride positive and a multiple of 16 bytes, 16-byte aligned bases,
// extents < 2^32.  Everything else stays on ew_tile_kernel.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "rc_kernel_common.cuh"

namespace rc {

constexpr int BK_T = 64;        // tile edge (elements)
constexpr int BK_STAGES = 4;    // source tiles in flight per CTA
constexpr int BK_OUTS = 2;      // output tile buffers
constexpr int BK_THREADS = 256;
constexpr uint32_t BK_TILE_BYTES = BK_T * BK_T * 8;
constexpr size_t BK_SMEM = (size_t)(BK_STAGES + BK_OUTS) * BK_TILE_BYTES + 64 + 1024;  // + barriers + alignment slack

// what the kernel needs beyond the two tensor maps: how a tile id becomes box coordinates
struct TmaTileDesc {
    uint32_t total_tiles;
    int nbatch;
    FastDiv div_ty, div_tx;
    FastDiv bdiv[3];
    int src_slot[4];  // coordinate slot (1..4) of X, batch 0, 1, 2 in the SOURCE map (slot 0 = Y)
    int dst_slot[4];  // coordinate slot (1..4) of Y, batch 0, 1, 2 in the OUTPUT map (slot 0 = X)
};

__device__ __forceinline__ uint32_t bk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bk_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bk_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bk_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bk_tma_load_5d(uint32_t dst_smem, const CUtensorMap *map, uint32_t bar, const int (&c)[5]) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst_smem), "l"(map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
__device__ __forceinline__ void bk_tma_store_5d(const CUtensorMap *map, uint32_t src_smem, const int (&c)[5]) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}
__device__ __forceinline__ void bk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bk_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bk_tile_coords(const TmaTileDesc &d, uint32_t tile, int (&cs)[5], int (&cd)[5]) {
    uint32_t t = tile, ty, tx;
    d.div_ty.divmod(t, t, ty);
    d.div_tx.divmod(t, t, tx);
#pragma unroll
    for (int i = 0; i < 5; ++i) { cs[i] = 0; cd[i] = 0; }
    cs[0] = (int)(ty * BK_T);             // source: Y is the inner dim
    cd[0] = (int)(tx * BK_T);             // output: X is the inner dim
    cs[d.src_slot[0]] = (int)(tx * BK_T);
    cd[d.dst_slot[0]] = (int)(ty * BK_T);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i >= d.nbatch) break;
        uint32_t q, r;
        d.bdiv[i].divmod(t, q, r);
        cs[d.src_slot[1 + i]] = (int)r;
        cd[d.dst_slot[1 + i]] = (int)r;
        t = q;
    }
}

template <int UNUSED = 0>  // a template so that only the translation unit that launches it carries the kernel
__global__ void __launch_bounds__(BK_THREADS, 1) ew_tile_tma_kernel(const __grid_constant__ CUtensorMap map_src,
                                                                   const __grid_constant__ CUtensorMap map_dst,
                                                                   const __grid_constant__ TmaTileDesc d) {
    extern __shared__ unsigned char bk_smem_raw[];
    // 1024-byte alignment: required by the 128-byte swizzle of the output tiles
    unsigned char *bk_smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(bk_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *out_tiles = reinterpret_cast<uint64_t *>(bk_smem);                                        // [OUTS][4][64][16]
    uint64_t *in_tiles = reinterpret_cast<uint64_t *>(bk_smem + (size_t)BK_OUTS * BK_TILE_BYTES);       // [STAGES][64][64]
    uint64_t *bars = reinterpret_cast<uint64_t *>(bk_smem + (size_t)(BK_OUTS + BK_STAGES) * BK_TILE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t first = blockIdx.x, step = gridDim.x, total = d.total_tiles;
    const uint32_t my_tiles = first < total ? (total - first + step - 1) / step : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < BK_STAGES; ++s) bk_mbar_init(bk_smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_load = [&](uint32_t k) {  // thread 0: k-th tile of this CTA into stage k % STAGES
        const int s = k % BK_STAGES;
        int cs[5], cd[5];
        bk_tile_coords(d, first + k * step, cs, cd);
        const uint32_t bar = bk_smem_u32(bars + s);
        bk_mbar_expect_tx(bar, BK_TILE_BYTES);
        bk_tma_load_5d(bk_smem_u32(in_tiles + (size_t)s * BK_T * BK_T), &map_src, bar, cs);
    };
    if (tid == 0) {
        for (uint32_t k = 0; k < (uint32_t)BK_STAGES && k < my_tiles; ++k) issue_load(k);
    }

    for (uint32_t k = 0; k < my_tiles; ++k) {
        const int s = k % BK_STAGES, ob = k % BK_OUTS;
        const uint32_t parity = (k / BK_STAGES) & 1u;
        bk_mbar_wait(bk_smem_u32(bars + s), parity);   // the source tile has landed
        if (tid == 0) bk_wait_read<BK_OUTS - 1>();     // the engine has read out[ob] of tile k - OUTS
        __syncthreads();
        const uint64_t *in = in_tiles + (size_t)s * BK_T * BK_T;
        ulonglong2 *out16 = reinterpret_cast<ulonglong2 *>(out_tiles + (size_t)ob * BK_T * BK_T);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int xp = warp + 8 * r;  // pair of x: 2 xp, 2 xp + 1; sub-tile q = xp / 8, chunk = xp % 8
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int y = lane + 32 * j;
                ulonglong2 v;
                v.x = in[(2 * xp) * BK_T + y];
                v.y = in[(2 * xp + 1) * BK_T + y];
                out16[(xp >> 3) * 512 + y * 8 + ((xp & 7) ^ (y & 7))] = v;
            }
        }
        bk_fence_async();   // generic-proxy writes of out[ob] before the engine reads them
        __syncthreads();    // ... and every read of in[s] is done
        if (tid == 0) {
            int cs[5], cd[5];
            bk_tile_coords(d, first + k * step, cs, cd);
            const uint32_t src = bk_smem_u32(out16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                bk_tma_store_5d(&map_dst, src + q * 8192, cd);
                cd[0] += 16;
            }
            bk_commit();
            if (k + BK_STAGES < my_tiles) issue_load(k + BK_STAGES);
        }
    }
    if (tid == 0) bk_wait_all();
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct TmaDim {
    uint64_t extent;
    int64_t stride;   // elements
    uint32_t box;
    int logical;      // 0 = the tiled outer dim (X for the source, Y for the output), 1.. = batch dims
};

// One 5-D map over 8-byte elements: dim 0 = the unit-stride axis (box `box0`), then `dims` sorted by stride, padded with
// extent-1 dims.  Returns false when the tensor cannot be described (the caller keeps the LDG / STG tile kernel).
inline bool tma_make_map(CUtensorMap *map, void *base, uint64_t extent0, uint32_t box0, std::vector<TmaDim> dims, int *slots,
                         CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = tma_encode_fn();
    if (!enc || dims.size() > 4) return false;
    std::stable_sort(dims.begin(), dims.end(), [](const TmaDim &a, const TmaDim &b) { return a.stride < b.stride; });
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t box[5], estr[5];
    gdim[0] = extent0;
    box[0] = box0;
    for (int i = 0; i < 5; ++i) estr[i] = 1;
    uint64_t prev_span = extent0 * 8;  // bytes covered so far: padding dims get a stride that is a multiple of 16
    for (int i = 0; i < 4; ++i) {
        if (i < (int)dims.size()) {
            const TmaDim &t = dims[i];
            if (t.stride <= 0 || (t.stride * 8) % 16 != 0 || t.extent == 0 || t.extent >= (1ull << 32)) return false;
            if ((uint64_t)t.stride * 8 >= (1ull << 40)) return false;
            gdim[i + 1] = t.extent;
            gstride[i] = (uint64_t)t.stride * 8;
            box[i + 1] = t.box;
            slots[t.logical] = i + 1;
            prev_span = std::max<uint64_t>(prev_span, gstride[i] * t.extent);
        } else {
            gdim[i + 1] = 1;
            gstride[i] = (prev_span + 15) & ~(uint64_t)15;
            box[i + 1] = 1;
        }
    }
    if (extent0 >= (1ull << 32) || reinterpret_cast<uintptr_t>(base) % 16 != 0) return false;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace rc
