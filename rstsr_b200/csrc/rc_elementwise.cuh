// rc_elementwise.cuh -- kernel family K1/K2: N-ary elementwise over strided operands.
//
// Replaces the CPU loops of rstsr-native-impl/src/cpu_rayon/op_with_func.rs:13-390 and
// cpu_rayon/assignment.rs:95-225 (same-index pairing; every operand already broadcast to one shape).
//
//  * ew_kernel<F, VEC>      flat index space over <= KMAXD merged dims.  VEC > 1: dim 0 is contiguous
//                           (stride 1, or 0 = splat) in every operand and moved as 16-byte packs;
//                           VEC == 1: arbitrary strides (negative / zero included).
//  * ew_tile_kernel<F>      operands disagree on the fastest axis (transposed copy, a + b^T):
//                           inputs whose unit stride lies on another axis Y are staged through a
//                           padded shared-memory tile so global reads run along Y and global writes
//                           along the output's axis X -- both sides stay coalesced.
//
// All kernels are HBM-bound: no tensor cores, grid sized from the element count, 64 B (f64) of loads
// in flight per thread and operand.
#pragma once
#include <type_traits>

#include "rc_kernel_common.cuh"

namespace rc {

enum OperandMode : int {
    MODE_MEM = 0,    // read through the operand's strides
    MODE_CONST = 1   // host scalar passed in the kernel parameters (`numa` / `numb` variants, fill)
};

constexpr int EW_BLOCK = 256;
constexpr int EW_UNROLL = 4;

template <class T>
struct EwConst {  // scalar operand slot; T may be any POD
    T v;
};

// ---------------------------------------------------------------------------------------------
// flat kernel.  F::NIN in {0 (fill), 1, 2}; F::apply(a[, b]) -> F::TO.
// ---------------------------------------------------------------------------------------------
template <class F, int VEC>
__global__ void __launch_bounds__(EW_BLOCK) ew_kernel(const EwDesc<3> d, typename F::TO *c,
                                                      const typename F::TA *a, const typename F::TB *b, int mode_a,
                                                      int mode_b, EwConst<typename F::TA> ka,
                                                      EwConst<typename F::TB> kb) {
    using TA = typename F::TA;
    using TB = typename F::TB;
    using TO = typename F::TO;
    const uint32_t first = blockIdx.x * (EW_BLOCK * EW_UNROLL) + threadIdx.x;

    Pack<TA, VEC> va[EW_UNROLL];
    Pack<TB, VEC> vb[EW_UNROLL];
    int64_t oc[EW_UNROLL];
    // issue every load of this thread before the first use: EW_UNROLL * 16 B in flight per operand
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        uint32_t idx = first + u * EW_BLOCK;
        if (idx < d.total) {
            int64_t off[3];
            ew_offsets<3>(d, idx, off);
            oc[u] = off[0];
            if (mode_a == MODE_MEM) {
                if (VEC > 1 && d.stride[1][0] == 0) {  // broadcast along the fastest axis
                    TA s = a[off[1]];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) va[u].v[j] = s;
                } else {
                    va[u] = ld_stream<TA, VEC>(a + off[1]);
                }
            }
            if (F::NIN > 1 && mode_b == MODE_MEM) {
                if (VEC > 1 && d.stride[2][0] == 0) {
                    TB s = b[off[2]];
#pragma unroll
                    for (int j = 0; j < VEC; ++j) vb[u].v[j] = s;
                } else {
                    vb[u] = ld_stream<TB, VEC>(b + off[2]);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        uint32_t idx = first + u * EW_BLOCK;
        if (idx < d.total) {
            Pack<TO, VEC> r;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                TA x = (mode_a == MODE_MEM) ? va[u].v[j] : ka.v;
                if constexpr (F::NIN > 1) {
                    TB y = (mode_b == MODE_MEM) ? vb[u].v[j] : kb.v;
                    r.v[j] = F::apply(x, y);
                } else {
                    r.v[j] = F::apply(x);
                }
            }
            st_stream<TO, VEC>(c + oc[u], r);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tile kernel.  X = canonical dim 0 (output stride 1), Y = the staged inputs' unit-stride dim.
// ---------------------------------------------------------------------------------------------
constexpr int TILE_X = 64;
constexpr int TILE_Y = 64;
constexpr int TILE_WARPS = 8;

enum TileMode : int { TILE_DIRECT = 0, TILE_STAGED = 1, TILE_CONST = 2 };

struct TileDesc {
    uint32_t nx, ny;         // extents of X and Y
    uint32_t tiles_x, tiles_y;
    int nbatch;              // remaining dims
    uint32_t total_tiles;
    FastDiv div_ty, div_tx;  // tile id -> (ty, tx, batch index)
    FastDiv bdiv[KMAXD];     // batch extents
    int64_t bstride[3][KMAXD];
    int64_t sx[3], sy[3];    // strides along X and Y per operand (c: sx = 1)
};

template <class F>
__global__ void __launch_bounds__(TILE_WARPS * 32) ew_tile_kernel(const TileDesc d, typename F::TO *c,
                                                                   const typename F::TA *a, const typename F::TB *b,
                                                                   int mode_a, int mode_b,
                                                                   EwConst<typename F::TA> ka,
                                                                   EwConst<typename F::TB> kb) {
    using TA = typename F::TA;
    using TB = typename F::TB;
    using TO = typename F::TO;
    constexpr int PITCH = TILE_Y + 1;  // odd pitch: column reads hit distinct banks (8-byte: per half-warp)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TA *sa = reinterpret_cast<TA *>(smem_raw);
    TB *sb = reinterpret_cast<TB *>(smem_raw + ((mode_a == TILE_STAGED) ? sizeof(TA) * TILE_X * PITCH : 0));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t t = blockIdx.x, ty, tx;
    d.div_ty.divmod(t, t, ty);
    d.div_tx.divmod(t, t, tx);
    int64_t base[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < KMAXD; ++i) {
        if (i >= d.nbatch) break;
        uint32_t q, r;
        d.bdiv[i].divmod(t, q, r);
#pragma unroll
        for (int k = 0; k < 3; ++k) base[k] += (int64_t)r * d.bstride[k][i];
        t = q;
    }
    const uint32_t x0 = tx * TILE_X, y0 = ty * TILE_Y;

    // phase 1: staged operands, lanes along Y (their contiguous axis)
    constexpr int RX = TILE_X / TILE_WARPS, CY = TILE_Y / 32;
    if (mode_a == TILE_STAGED) {
        TA reg[RX][CY];
#pragma unroll
        for (int r = 0; r < RX; ++r) {
            uint32_t x = x0 + warp + r * TILE_WARPS;
#pragma unroll
            for (int j = 0; j < CY; ++j) {
                uint32_t y = y0 + lane + 32 * j;
                if (x < d.nx && y < d.ny) reg[r][j] = __ldcs(a + base[1] + (int64_t)x * d.sx[1] + y);
            }
        }
#pragma unroll
        for (int r = 0; r < RX; ++r)
#pragma unroll
            for (int j = 0; j < CY; ++j) sa[(warp + r * TILE_WARPS) * PITCH + lane + 32 * j] = reg[r][j];
    }
    if (F::NIN > 1 && mode_b == TILE_STAGED) {
        TB reg[RX][CY];
#pragma unroll
        for (int r = 0; r < RX; ++r) {
            uint32_t x = x0 + warp + r * TILE_WARPS;
#pragma unroll
            for (int j = 0; j < CY; ++j) {
                uint32_t y = y0 + lane + 32 * j;
                if (x < d.nx && y < d.ny) reg[r][j] = __ldcs(b + base[2] + (int64_t)x * d.sx[2] + y);
            }
        }
#pragma unroll
        for (int r = 0; r < RX; ++r)
#pragma unroll
            for (int j = 0; j < CY; ++j) sb[(warp + r * TILE_WARPS) * PITCH + lane + 32 * j] = reg[r][j];
    }
    __syncthreads();

    // phase 2: lanes along X (the output's contiguous axis)
    constexpr int RY = TILE_Y / TILE_WARPS, CX = TILE_X / 32;
    TA xa[RY][CX];
    TB xb[RY][CX];
#pragma unroll
    for (int r = 0; r < RY; ++r) {
        uint32_t y = y0 + warp + r * TILE_WARPS;
#pragma unroll
        for (int j = 0; j < CX; ++j) {
            uint32_t x = x0 + lane + 32 * j;
            if (x < d.nx && y < d.ny) {
                if (mode_a == TILE_DIRECT) xa[r][j] = __ldcs(a + base[1] + (int64_t)x * d.sx[1] + (int64_t)y * d.sy[1]);
                if (F::NIN > 1 && mode_b == TILE_DIRECT)
                    xb[r][j] = __ldcs(b + base[2] + (int64_t)x * d.sx[2] + (int64_t)y * d.sy[2]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RY; ++r) {
        const int yl = warp + r * TILE_WARPS;
        uint32_t y = y0 + yl;
#pragma unroll
        for (int j = 0; j < CX; ++j) {
            const int xl = lane + 32 * j;
            uint32_t x = x0 + xl;
            if (x < d.nx && y < d.ny) {
                TA va = (mode_a == TILE_DIRECT) ? xa[r][j] : (mode_a == TILE_STAGED ? sa[xl * PITCH + yl] : ka.v);
                TO out;
                if constexpr (F::NIN > 1) {
                    TB vb = (mode_b == TILE_DIRECT) ? xb[r][j] : (mode_b == TILE_STAGED ? sb[xl * PITCH + yl] : kb.v);
                    out = F::apply(va, vb);
                } else {
                    out = F::apply(va);
                }
                __stcs(c + base[0] + (int64_t)y * d.sy[0] + x, out);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
struct EwArgs {
    void *c = nullptr;
    const void *a = nullptr, *b = nullptr;
    bool a_const = false, b_const = false;  // operand is a host scalar
    const void *a_host = nullptr, *b_host = nullptr;
};

// Splits a canonical problem until it fits one launch (<= KMAXD dims, < 2^31 items); fn(part).
template <class Fn>
void ew_for_each_part(const CanonEw &c, Fn &&fn) {
    if (c.empty) return;
    int64_t total = c.total();
    if (c.ndim <= KMAXD && total <= kMaxItemsPerLaunch) {
        fn(c);
        return;
    }
    int last = c.ndim - 1;
    int64_t inner = total / c.shape[last];
    // chunk of the outermost dim per part
    int64_t chunk = (c.ndim > KMAXD) ? 1 : std::max<int64_t>(1, kMaxItemsPerLaunch / inner);
    if (c.ndim == 1) chunk = (1ll << 30);  // keeps 16-byte alignment of every part
    for (int64_t s = 0; s < c.shape[last]; s += chunk) {
        CanonEw p = c;
        int64_t n = std::min(chunk, c.shape[last] - s);
        for (int k = 0; k < c.nops; ++k) p.base[k] += s * c.stride[k][last];
        if (n == 1 && c.ndim > 1) {
            p.ndim = c.ndim - 1;
            p.shape.pop_back();
            for (int k = 0; k < c.nops; ++k) p.stride[k].pop_back();
        } else {
            p.shape[last] = n;
        }
        ew_for_each_part(p, fn);
    }
}

template <class F, bool ALLOW_TILE, bool ALLOW_VEC>
void ew_launch_part(rc_device *dev, const CanonEw &c, const EwArgs &args) {
    using TA = typename F::TA;
    using TB = typename F::TB;
    using TO = typename F::TO;
    constexpr int NIN = F::NIN;
    constexpr size_t maxsz = sizeof(TO) > sizeof(TA) ? (sizeof(TO) > sizeof(TB) ? sizeof(TO) : sizeof(TB))
                                                     : (sizeof(TA) > sizeof(TB) ? sizeof(TA) : sizeof(TB));
    constexpr int V = ALLOW_VEC ? (int)(16 / maxsz) : 1;

    TO *pc = static_cast<TO *>(args.c) + c.base[0];
    const TA *pa = nullptr;
    const TB *pb = nullptr;
    EwConst<TA> ka;
    EwConst<TB> kb;
    std::memset(&ka, 0, sizeof(ka));
    std::memset(&kb, 0, sizeof(kb));
    int mode_a = MODE_CONST, mode_b = MODE_CONST;
    // operand slots in the canonical form: 0 = c, then the memory operands in order (a, b)
    int slot = 1, slot_a = -1, slot_b = -1;
    if (NIN >= 1) {
        if (args.a_const) std::memcpy(&ka.v, args.a_host, sizeof(TA));
        else { slot_a = slot++; pa = static_cast<const TA *>(args.a) + c.base[slot_a]; mode_a = MODE_MEM; }
    }
    if (NIN >= 2) {
        if (args.b_const) std::memcpy(&kb.v, args.b_host, sizeof(TB));
        else { slot_b = slot++; pb = static_cast<const TB *>(args.b) + c.base[slot_b]; mode_b = MODE_MEM; }
    }
    if (NIN == 0 && args.a_host) std::memcpy(&ka.v, args.a_host, sizeof(TA));  // fill value

    auto stride_of = [&](int s, int i) -> int64_t { return s < 0 ? 0 : c.stride[s][i]; };

    // ---- can dim 0 move as 16-byte packs? ----
    bool vec_ok = V > 1 && c.stride[0][0] == 1 && (c.shape[0] % V == 0) &&
                  (reinterpret_cast<uintptr_t>(pc) % (V * sizeof(TO)) == 0);
    auto vec_operand_ok = [&](int s, const void *p, size_t esz) {
        if (s < 0) return true;
        int64_t s0 = c.stride[s][0];
        if (s0 == 0) return true;  // splat
        if (s0 != 1) return false;
        if (reinterpret_cast<uintptr_t>(p) % (V * esz) != 0) return false;
        for (int i = 1; i < c.ndim; ++i)
            if (c.stride[s][i] % V != 0) return false;
        return true;
    };
    if (vec_ok) {
        for (int i = 1; i < c.ndim; ++i) vec_ok = vec_ok && (c.stride[0][i] % V == 0);
        vec_ok = vec_ok && vec_operand_ok(slot_a, pa, sizeof(TA)) && vec_operand_ok(slot_b, pb, sizeof(TB));
    }

    // ---- tile path: output contiguous on dim 0, some input contiguous on another dim ----
    if constexpr (ALLOW_TILE && NIN >= 1)
    if (!vec_ok && c.ndim >= 2 && c.stride[0][0] == 1 && c.shape[0] >= 16) {
        int ydim = -1;
        auto unit_dim = [&](int s) {
            if (s < 0) return -1;
            if (c.stride[s][0] == 0 || c.stride[s][0] == 1) return -1;  // already fine along X
            for (int i = 1; i < c.ndim; ++i)
                if (c.stride[s][i] == 1 && c.shape[i] >= 16) return i;
            return -1;
        };
        int ya = unit_dim(slot_a), yb = unit_dim(slot_b);
        ydim = ya >= 0 ? ya : yb;
        if (ydim >= 0) {
            TileDesc t;
            std::memset(&t, 0, sizeof(t));
            t.nx = (uint32_t)c.shape[0];
            t.ny = (uint32_t)c.shape[ydim];
            t.tiles_x = (t.nx + TILE_X - 1) / TILE_X;
            t.tiles_y = (t.ny + TILE_Y - 1) / TILE_Y;
            t.div_tx = FastDiv(t.tiles_x);
            t.div_ty = FastDiv(t.tiles_y);
            int64_t nb = 1;
            int bi = 0;
            const int slots[3] = {0, slot_a, slot_b};
            for (int i = 1; i < c.ndim; ++i) {
                if (i == ydim) continue;
                t.bdiv[bi] = FastDiv((uint32_t)c.shape[i]);
                for (int k = 0; k < 3; ++k) t.bstride[k][bi] = stride_of(slots[k], i);
                nb *= c.shape[i];
                ++bi;
            }
            t.nbatch = bi;
            for (int k = 0; k < 3; ++k) {
                t.sx[k] = stride_of(slots[k], 0);
                t.sy[k] = stride_of(slots[k], ydim);
            }
            int64_t total_tiles = (int64_t)t.tiles_x * t.tiles_y * nb;
            if (total_tiles < (1ll << 31)) {
                t.total_tiles = (uint32_t)total_tiles;
                int tm_a = TILE_CONST, tm_b = TILE_CONST;
                if (slot_a >= 0) tm_a = (c.stride[slot_a][ydim] == 1 && c.stride[slot_a][0] > 1 && ya == ydim) ? TILE_STAGED : TILE_DIRECT;
                if (slot_b >= 0) tm_b = (c.stride[slot_b][ydim] == 1 && c.stride[slot_b][0] > 1 && yb == ydim) ? TILE_STAGED : TILE_DIRECT;
                size_t smem = 0;
                if (tm_a == TILE_STAGED) smem += sizeof(TA) * TILE_X * (TILE_Y + 1);
                if (tm_b == TILE_STAGED) smem += sizeof(TB) * TILE_X * (TILE_Y + 1);
                if (smem > 48 * 1024)  // opt in to > 48 KB dynamic shared memory (per device)
                    RC_CUDA(cudaFuncSetAttribute(ew_tile_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem));
                ew_tile_kernel<F><<<t.total_tiles, TILE_WARPS * 32, smem, dev->stream>>>(t, pc, pa, pb, tm_a, tm_b, ka, kb);
                after_launch(dev, "ew_tile_kernel");
                return;
            }
        }
    }

    EwDesc<3> d;
    std::memset(&d, 0, sizeof(d));
    d.ndim = c.ndim;
    const int slots[3] = {0, slot_a, slot_b};
    const int vec = vec_ok ? V : 1;
    int64_t items = 1;
    for (int i = 0; i < c.ndim; ++i) {
        int64_t n = (i == 0) ? c.shape[0] / vec : c.shape[i];
        d.div[i] = FastDiv((uint32_t)n);
        items *= n;
        for (int k = 0; k < 3; ++k) {
            int64_t s = stride_of(slots[k], i);
            d.stride[k][i] = (i == 0) ? s * vec : s;
        }
    }
    d.total = (uint32_t)items;
    uint32_t grid = (uint32_t)((items + EW_BLOCK * EW_UNROLL - 1) / (EW_BLOCK * EW_UNROLL));
    if constexpr (V > 1) {
        if (vec_ok) {
            ew_kernel<F, V><<<grid, EW_BLOCK, 0, dev->stream>>>(d, pc, pa, pb, mode_a, mode_b, ka, kb);
            after_launch(dev, "ew_kernel");
            return;
        }
    }
    ew_kernel<F, 1><<<grid, EW_BLOCK, 0, dev->stream>>>(d, pc, pa, pb, mode_a, mode_b, ka, kb);
    after_launch(dev, "ew_kernel");
}

template <class F, bool ALLOW_TILE = true, bool ALLOW_VEC = true>
void ew_launch(rc_device *dev, const CanonEw &c, const EwArgs &args) {
    ew_for_each_part(c, [&](const CanonEw &p) { ew_launch_part<F, ALLOW_TILE, ALLOW_VEC>(dev, p, args); });
}

}  // namespace rc
