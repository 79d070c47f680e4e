// rc_elementwise.cuh -- kernel family K1/K2: N-ary elementwise over strided operands.
//
// Replaces the CPU loops of rstsr-native-impl/src/cpu_rayon/op_with_func.rs:13-390 and
// cpu_rayon/assignment.rs:95-225 (same-index pairing; every operand already broadcast to one shape).
//
//  * ew_kernel<F, VEC, ND>  flat index space over <= KMAXD merged dims.  VEC > 1: dim 0 is contiguous
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

template <class T> struct FIdentity;  // rc_functors.cuh (the bit-moving copy the TMA-engine tile kernel serves)

enum OperandMode : int {
    MODE_MEM = 0,    // read through the operand's strides
    MODE_CONST = 1,  // host scalar passed in the kernel parameters (`numa` / `numb` variants, fill)
    MODE_SPLAT = 2   // vector kernel only: stride 0 along dim 0 -> one element broadcast over the pack
};

constexpr int EW_BLOCK = 256;
constexpr int EW_UNROLL = 4;

template <class T>
struct EwConst {  // scalar operand slot; T may be any POD
    T v;
};

// Run-time parameters of a functor (isclose: rtol / atol / equal_nan): a functor that declares `using Params = ...`
// receives them as the last argument of apply(); every other functor gets an empty struct the compiler drops.
struct EwNoParams {};
template <class F, class = void> struct ew_params { using type = EwNoParams; };
template <class F> struct ew_params<F, std::void_t<typename F::Params>> { using type = typename F::Params; };
template <class F> using ew_params_t = typename ew_params<F>::type;
template <class F>
__device__ __forceinline__ typename F::TO ew_apply(typename F::TA x, typename F::TB y, const ew_params_t<F> &prm) {
    if constexpr (std::is_same<ew_params_t<F>, EwNoParams>::value) return F::apply(x, y);
    else return F::apply(x, y, prm);
}

// offsets of work item `idx` in the three operands.  ND = 1, 2: compile-time rank (no loop, at most one
// division); ND = 0: runtime rank d.ndim <= KMAXD.
template <int ND>
__device__ __forceinline__ void ew_item_offsets(const EwDesc<3> &d, uint32_t idx, int64_t &oc, int64_t &oa,
                                                int64_t &ob) {
    if constexpr (ND == 1) {
        oc = (int64_t)idx * d.stride[0][0];
        oa = (int64_t)idx * d.stride[1][0];
        ob = (int64_t)idx * d.stride[2][0];
    } else if constexpr (ND == 2) {
        uint32_t q, r;
        d.div[0].divmod(idx, q, r);
        oc = (int64_t)r * d.stride[0][0] + (int64_t)q * d.stride[0][1];
        oa = (int64_t)r * d.stride[1][0] + (int64_t)q * d.stride[1][1];
        ob = (int64_t)r * d.stride[2][0] + (int64_t)q * d.stride[2][1];
    } else {
        int64_t off[3];
        ew_offsets<3>(d, idx, off);
        oc = off[0];
        oa = off[1];
        ob = off[2];
    }
}

// ---------------------------------------------------------------------------------------------
// flat kernel.  F::NIN in {0 (fill), 1, 2}; F::apply(a[, b]) -> F::TO.  One-shot grid (measured on B200:
// one-shot grids reach 6.8-7.0 TB/s on copy/add, persistent grid-stride loops 5.8-6.5 TB/s).
// ---------------------------------------------------------------------------------------------
template <class F, int VEC, int ND, int UN = EW_UNROLL>
__global__ void __launch_bounds__(EW_BLOCK) ew_kernel(const __grid_constant__ EwDesc<3> d, typename F::TO *c, const typename F::TA *a,
                                                      const typename F::TB *b, int mode_a, int mode_b,
                                                      EwConst<typename F::TA> ka, EwConst<typename F::TB> kb,
        ew_params_t<F> prm) {
    using TA = typename F::TA;
    using TB = typename F::TB;
    using TO = typename F::TO;
    const uint32_t first = blockIdx.x * (EW_BLOCK * UN) + threadIdx.x;

    // Straight-line code for every operand mode: packs start out as the constant, a predicated in-place
    // LDG overwrites them for memory operands, splat operands get a predicated scalar LDG and are selected
    // at the point of use.  All loads of the thread are issued before the first use.
    Pack<TA, VEC> va[UN];
    Pack<TB, VEC> vb[UN];
    Pack<TA, 1> sa[UN];
    Pack<TB, 1> sb[UN];
    int64_t oc[UN];
    bool ok[UN];
    const bool mem_a = (F::NIN >= 1) && mode_a == MODE_MEM, spl_a = (F::NIN >= 1) && mode_a == MODE_SPLAT;
    const bool mem_b = (F::NIN >= 2) && mode_b == MODE_MEM, spl_b = (F::NIN >= 2) && mode_b == MODE_SPLAT;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        const uint32_t idx = first + u * EW_BLOCK;
        ok[u] = idx < d.total;
        int64_t oa, ob;
        ew_item_offsets<ND>(d, idx, oc[u], oa, ob);
#pragma unroll
        for (int j = 0; j < VEC; ++j) va[u].v[j] = ka.v;
        sa[u].v[0] = ka.v;
        if constexpr (F::NIN >= 1) {
            ld_stream_pred<TA, VEC>(va[u], a + oa, ok[u] && mem_a);
            if constexpr (VEC > 1) ld_stream_pred<TA, 1>(sa[u], a + oa, ok[u] && spl_a);
        }
        if constexpr (F::NIN >= 2) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) vb[u].v[j] = kb.v;
            sb[u].v[0] = kb.v;
            ld_stream_pred<TB, VEC>(vb[u], b + ob, ok[u] && mem_b);
            if constexpr (VEC > 1) ld_stream_pred<TB, 1>(sb[u], b + ob, ok[u] && spl_b);
        }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        if (ok[u]) {
            Pack<TO, VEC> r;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const TA x = (VEC > 1 && spl_a) ? sa[u].v[0] : va[u].v[j];
                if constexpr (F::NIN > 1) {
                    const TB y = (VEC > 1 && spl_b) ? sb[u].v[0] : vb[u].v[j];
                    r.v[j] = ew_apply<F>(x, y, prm);
                } else {
                    r.v[j] = F::apply(x);
                }
            }
            st_stream<TO, VEC>(c + oc[u], r);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// two-dim kernel: dim 0 (packs) x dim 1 (rows).  A CTA owns one 1024-pack chunk of dim 0 and walks R rows.
// Operands that are broadcast over dim 1 (stride 0: `(8192,8192) + (8192,)`, `a * v`) are loaded ONCE per
// CTA and stay in registers for all R rows (R is 1 by default, see the launcher: one-shot grids measured
// fastest).  No per-item division at all (one fast division per CTA).
// ---------------------------------------------------------------------------------------------
struct EwRowsDesc {
    uint32_t n0;              // packs along dim 0
    uint32_t n1;              // rows
    uint32_t rows_per_cta;    // R
    FastDiv div_chunks;       // chunks per row
    int64_t s0[3], s1[3];     // element strides per pack step / per row
};

template <class F, int VEC>
__global__ void __launch_bounds__(EW_BLOCK) ew_rows_kernel(const __grid_constant__ EwRowsDesc d, typename F::TO *c,
                                                           const typename F::TA *a, const typename F::TB *b,
                                                           int mode_a, int mode_b, EwConst<typename F::TA> ka,
                                                           EwConst<typename F::TB> kb, ew_params_t<F> prm) {
    using TA = typename F::TA;
    using TB = typename F::TB;
    using TO = typename F::TO;
    uint32_t rg, chunk;
    d.div_chunks.divmod(blockIdx.x, rg, chunk);
    const uint32_t q0 = rg * d.rows_per_cta;
    const uint32_t rows = min(d.rows_per_cta, d.n1 - q0);
    const uint32_t first = chunk * (EW_BLOCK * EW_UNROLL) + threadIdx.x;

    // operand classes, uniform over the CTA: streamed per row / kept in registers (no dim-1 stride) / constant.
    // (operands broadcast along dim 0 -- MODE_SPLAT -- never reach this kernel unless they are also kept)
    const bool keep_a = (F::NIN >= 1) && (mode_a != MODE_CONST) && d.s1[1] == 0;
    const bool keep_b = (F::NIN >= 2) && (mode_b != MODE_CONST) && d.s1[2] == 0;
    const bool stream_a = (F::NIN >= 1) && mode_a == MODE_MEM && !keep_a;
    const bool stream_b = (F::NIN >= 2) && mode_b == MODE_MEM && !keep_b;

    // one base pointer per operand; item u of the thread sits u * EW_BLOCK packs further along dim 0
    TO *pc = c + ((int64_t)first * d.s0[0] + (int64_t)q0 * d.s1[0]);
    const TA *pa = a + ((int64_t)first * d.s0[1] + (int64_t)q0 * d.s1[1]);
    const TB *pb = b + ((int64_t)first * d.s0[2] + (int64_t)q0 * d.s1[2]);
    const int64_t step_c = (int64_t)EW_BLOCK * d.s0[0], step_a = (int64_t)EW_BLOCK * d.s0[1],
                  step_b = (int64_t)EW_BLOCK * d.s0[2];

    bool ok[EW_UNROLL];
    Pack<TA, VEC> va[EW_UNROLL];  // packs start out as the constant; kept operands are loaded once here,
    Pack<TB, VEC> vb[EW_UNROLL];  // streamed ones once per row, all through predicated in-place LDGs
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        ok[u] = first + u * EW_BLOCK < d.n0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) va[u].v[j] = ka.v;
        if constexpr (F::NIN >= 1) {
            if (mode_a == MODE_SPLAT) {  // one value per (kept) row chunk: broadcast over the pack
                Pack<TA, 1> s;
                s.v[0] = ka.v;
                ld_stream_pred<TA, 1>(s, pa + u * step_a, ok[u]);
#pragma unroll
                for (int j = 0; j < VEC; ++j) va[u].v[j] = s.v[0];
            } else {
                ld_stream_pred<TA, VEC>(va[u], pa + u * step_a, ok[u] && keep_a);
            }
        }
        if constexpr (F::NIN >= 2) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) vb[u].v[j] = kb.v;
            if (mode_b == MODE_SPLAT) {
                Pack<TB, 1> s;
                s.v[0] = kb.v;
                ld_stream_pred<TB, 1>(s, pb + u * step_b, ok[u]);
#pragma unroll
                for (int j = 0; j < VEC; ++j) vb[u].v[j] = s.v[0];
            } else {
                ld_stream_pred<TB, VEC>(vb[u], pb + u * step_b, ok[u] && keep_b);
            }
        }
    }

    for (uint32_t row = 0; row < rows; ++row) {
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) {
            if constexpr (F::NIN >= 1) ld_stream_pred<TA, VEC>(va[u], pa + u * step_a, ok[u] && stream_a);
            if constexpr (F::NIN >= 2) ld_stream_pred<TB, VEC>(vb[u], pb + u * step_b, ok[u] && stream_b);
        }
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) {
            if (ok[u]) {
                Pack<TO, VEC> r;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if constexpr (F::NIN > 1) r.v[j] = ew_apply<F>(va[u].v[j], vb[u].v[j], prm);
                    else r.v[j] = F::apply(va[u].v[j]);
                }
                st_stream<TO, VEC>(pc + u * step_c, r);
            }
        }
        pc += d.s1[0];
        pa += d.s1[1];
        pb += d.s1[2];
    }
}

// ---------------------------------------------------------------------------------------------
// outer kernel: WRITE-ONLY broadcast ops c[i, j] = f(u[i], v[j]) (outer sum / product, `a[:, None] * b[None, :]`).
// One operand is constant along the rows (kept in registers: packs of v, or a host scalar), the other is one scalar per
// row (stride 0 along dim 0).  A CTA owns a 1024-pack chunk of dim 0 and walks R rows: the kept packs are loaded once,
// the row scalar of row r + 1 is fetched while row r is computed and stored, so the loop body is FADD/FMUL + STG only.
// (The flat kernel spent 98 instructions per 16-byte pack on this shape -- every operand mode materialised per slot --
// and was issue-bound at 4.4 TB/s, profiles/r01_instruction_mix.md.)
// SPL = 1: a is the row scalar, b is kept; SPL = 2: b is the row scalar, a is kept.
// ---------------------------------------------------------------------------------------------
constexpr int EW_OUTER_ROWS = 8;

template <class F, int VEC, int SPL>
__global__ void __launch_bounds__(EW_BLOCK) ew_outer_kernel(const __grid_constant__ EwRowsDesc d, typename F::TO *c,
                                                            const typename F::TA *a, const typename F::TB *b,
                                                            int mode_kept, EwConst<typename F::TA> ka,
                                                            EwConst<typename F::TB> kb, ew_params_t<F> prm) {
    using TA = typename F::TA;
    using TB = typename F::TB;
    using TO = typename F::TO;
    using TK = typename std::conditional<SPL == 1, TB, TA>::type;  // kept operand
    using TS = typename std::conditional<SPL == 1, TA, TB>::type;  // row scalar
    constexpr int KS = SPL == 1 ? 2 : 1, SS = SPL == 1 ? 1 : 2;     // their slots in the descriptor
    uint32_t rg, chunk;
    d.div_chunks.divmod(blockIdx.x, rg, chunk);
    const uint32_t q0 = rg * d.rows_per_cta;
    const uint32_t rows = min(d.rows_per_cta, d.n1 - q0);
    const uint32_t first = chunk * (EW_BLOCK * EW_UNROLL) + threadIdx.x;

    TO *pc = c + ((int64_t)first * d.s0[0] + (int64_t)q0 * d.s1[0]);
    const int64_t step_c = (int64_t)EW_BLOCK * d.s0[0], step_k = (int64_t)EW_BLOCK * d.s0[KS];
    const TK *pk;
    const TS *ps;
    TK kconst;
    if constexpr (SPL == 1) { pk = b + (int64_t)first * d.s0[2]; ps = a + (int64_t)q0 * d.s1[1]; kconst = kb.v; }
    else                    { pk = a + (int64_t)first * d.s0[1]; ps = b + (int64_t)q0 * d.s1[2]; kconst = ka.v; }

    bool ok[EW_UNROLL];
    Pack<TK, VEC> vk[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        ok[u] = first + u * EW_BLOCK < d.n0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) vk[u].v[j] = kconst;
        ld_stream_pred<TK, VEC>(vk[u], pk + u * step_k, ok[u] && mode_kept == MODE_MEM);
    }
    const int64_t ss = d.s1[SS];
    TS cur = *ps;
    for (uint32_t row = 0; row < rows; ++row) {
        TS nxt = cur;
        if (row + 1 < rows) nxt = ps[ss];  // every thread of the CTA reads the same word: one broadcast transaction
#pragma unroll
        for (int u = 0; u < EW_UNROLL; ++u) {
            if (ok[u]) {
                Pack<TO, VEC> r;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if constexpr (SPL == 1) r.v[j] = ew_apply<F>(cur, vk[u].v[j], prm);
                    else r.v[j] = ew_apply<F>(vk[u].v[j], cur, prm);
                }
                st_stream<TO, VEC>(pc + u * step_c, r);
            }
        }
        pc += d.s1[0];
        ps += ss;
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------
// tile kernel.  X = canonical dim 0 (output stride 1), Y = the staged inputs' unit-stride dim.
// ---------------------------------------------------------------------------------------------
// Smallest extents that take the tile kernel (tuning knobs RC_TILE_MIN_X / RC_TILE_MIN_Y for experiments):
//   X = the output's contiguous axis: below it the flat kernel writes short rows just as well,
//   Y = the staged operand's contiguous axis: reading it with the flat kernel is sector-granular however short it is.
// Measured (scripts/probe_smalldim.py, k = short extent): output rows shorter than 128 bytes are written as well by
// the flat kernel (f32 k = 16..24: 3.7 vs 2.5 TB/s), so X >= max(16, 128 / itemsize); a staged axis of 48 bytes or more
// pays for the tile (f64 k = 8 / 12: 1.6 -> 2.5 / 1.1 -> 3.6 TB/s, f32 k = 12: 1.1 -> 2.0), so Y >= 48 / itemsize.
inline int tile_min_x(size_t esz) {
    static int v = [] { const char *e = getenv("RC_TILE_MIN_X"); int x = e ? atoi(e) : 0; return x >= 2 ? x : 0; }();
    return v ? v : std::max<int>(16, (int)(128 / esz));
}
inline int tile_min_y(size_t esz) {
    static int v = [] { const char *e = getenv("RC_TILE_MIN_Y"); int x = e ? atoi(e) : 0; return x >= 2 ? x : 0; }();
    return v ? v : std::max<int>(2, (int)((48 + esz - 1) / esz));
}
// rows per CTA of ew_rows_kernel when an operand is broadcast over the rows (RC_ROWS_PER_CTA, experiments; default 1)
inline int ew_rows_per_cta() {
    static const int v = [] { const char *e = getenv("RC_ROWS_PER_CTA"); int x = e ? atoi(e) : 1; return x >= 1 ? x : 1; }();
    return v;
}
// RC_TILE_RECT=0 switches the rectangular-tile kernel off (experiments: the square kernel then takes those shapes)
inline bool tile_rect_enabled() {
    static bool v = [] { const char *e = getenv("RC_TILE_RECT"); return !(e && e[0] == '0'); }();
    return v;
}
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
__global__ void __launch_bounds__(TILE_WARPS * 32) ew_tile_kernel(const __grid_constant__ TileDesc d, typename F::TO *c,
                                                                   const typename F::TA *a, const typename F::TB *b,
                                                                   int mode_a, int mode_b,
                                                                   EwConst<typename F::TA> ka,
                                                                   EwConst<typename F::TB> kb, ew_params_t<F> prm) {
    using TA = typename F::TA;
    using TB = typename F::TB;
    using TO = typename F::TO;
    static_assert(TILE_X == TILE_Y, "slot mapping below assumes a square tile");
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
    const uint32_t remx = d.nx - x0, remy = d.ny - y0;  // valid extent of this tile

    // Slot (r, j) of a thread has tile coordinates (u, v) = (warp + 8 r, lane + 32 j).  A STAGED operand is read
    // with (x, y) = (u, v): lanes run along Y, its contiguous axis.  A DIRECT operand and the output use
    // (y, x) = (u, v): lanes run along X.  One register array per operand serves either role, every load is a
    // predicated in-place LDG, and all loads of both operands are issued before anything is consumed.
    constexpr int R = TILE_X / TILE_WARPS, C = TILE_Y / 32;
    const bool stg_a = (F::NIN >= 1) && mode_a == TILE_STAGED, dir_a = (F::NIN >= 1) && mode_a == TILE_DIRECT;
    const bool stg_b = (F::NIN >= 2) && mode_b == TILE_STAGED, dir_b = (F::NIN >= 2) && mode_b == TILE_DIRECT;
    // element offset of slot (u, v): (p0 + u) * su + (q0 + v) * sv
    const int64_t su_a = stg_a ? d.sx[1] : d.sy[1], sv_a = stg_a ? d.sy[1] : d.sx[1];
    const int64_t su_b = stg_b ? d.sx[2] : d.sy[2], sv_b = stg_b ? d.sy[2] : d.sx[2];
    const TA *pa = a + base[1] + (int64_t)(stg_a ? x0 : y0) * su_a + (int64_t)(stg_a ? y0 : x0) * sv_a;
    const TB *pb = b + base[2] + (int64_t)(stg_b ? x0 : y0) * su_b + (int64_t)(stg_b ? y0 : x0) * sv_b;
    const uint32_t lim_u_a = stg_a ? remx : remy, lim_v_a = stg_a ? remy : remx;
    const uint32_t lim_u_b = stg_b ? remx : remy, lim_v_b = stg_b ? remy : remx;

    Pack<TA, 1> ra[R][C];
    Pack<TB, 1> rb[R][C];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t u = warp + r * TILE_WARPS;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const uint32_t v = lane + 32 * j;
            ra[r][j].v[0] = ka.v;
            if constexpr (F::NIN >= 1)
                ld_stream_pred<TA, 1>(ra[r][j], pa + (int64_t)u * su_a + (int64_t)v * sv_a,
                                      (stg_a || dir_a) && u < lim_u_a && v < lim_v_a);
            if constexpr (F::NIN >= 2) {
                rb[r][j].v[0] = kb.v;
                ld_stream_pred<TB, 1>(rb[r][j], pb + (int64_t)u * su_b + (int64_t)v * sv_b,
                                      (stg_b || dir_b) && u < lim_u_b && v < lim_v_b);
            }
        }
    }
    if (stg_a) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < C; ++j) sa[(warp + r * TILE_WARPS) * PITCH + lane + 32 * j] = ra[r][j].v[0];
    }
    if (stg_b) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < C; ++j) sb[(warp + r * TILE_WARPS) * PITCH + lane + 32 * j] = rb[r][j].v[0];
    }
    __syncthreads();

    // phase 2: slot (u, v) = (y, x); lanes along X, the output's contiguous axis
    TO *pc = c + base[0] + (int64_t)y0 * d.sy[0] + x0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t u = warp + r * TILE_WARPS;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const uint32_t v = lane + 32 * j;
            if (u < remy && v < remx) {
                const TA va = stg_a ? sa[v * PITCH + u] : ra[r][j].v[0];
                TO out;
                if constexpr (F::NIN > 1) {
                    const TB vb = stg_b ? sb[v * PITCH + u] : rb[r][j].v[0];
                    out = ew_apply<F>(va, vb, prm);
                } else {
                    out = F::apply(va);
                }
                Pack<TO, 1> po;
                po.v[0] = out;
                st_stream<TO, 1>(pc + (int64_t)u * d.sy[0] + v, po);
            }
        }
    }
}

}  // namespace rc
#include "rc_tile_bulk.cuh"
#include "rc_tile_narrow.cuh"
#include "rc_tile_wide.cuh"
#include "rc_tile_short.cuh"
namespace rc {

template <class F> struct is_word_copy : std::false_type {};
template <> struct is_word_copy<FIdentity<uint8_t>> : std::true_type {};
template <> struct is_word_copy<FIdentity<uint16_t>> : std::true_type {};
template <> struct is_word_copy<FIdentity<uint32_t>> : std::true_type {};
template <> struct is_word_copy<FIdentity<uint64_t>> : std::true_type {};

// RC_TILE_SHORT=0 switches ew_tile_short_kernel off; RC_SHORT_MAX_K<S> = longest short extent of S-byte elements it takes
inline bool tile_short_enabled() {
    static bool v = [] { const char *e = getenv("RC_TILE_SHORT"); return !(e && e[0] == '0'); }();
    return v;
}
inline int tile_short_max_k(size_t esz) {
    auto knob = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    // above these the tile no longer fits one trip of loads and the partly filled square / word tile is faster
    // (profiles/r02_tile_short.md)
    static const int m1 = knob("RC_SHORT_MAX_K1", 64), m2 = knob("RC_SHORT_MAX_K2", 48),
                     m4 = knob("RC_SHORT_MAX_K4", 40), m8 = knob("RC_SHORT_MAX_K8", 32);
    return std::min<int>(SHORT_MAX_K, esz == 1 ? m1 : esz == 2 ? m2 : esz == 4 ? m4 : m8);
}

// kind 1: short Y, flat source (de-interleave); kind 2: short X, flat output (interleave).  `t` is the TileDesc of a staged
// one-operand copy (slot 0 = output, slot 1 = source).
template <int S, class TD>
bool launch_tile_short(rc_device *dev, const TD &t, int kind, void *pc, const void *pa) {
    ShortDesc d;
    std::memset(&d, 0, sizeof(d));
    const bool deint = kind == 1;
    d.k = deint ? t.ny : t.nx;
    d.n = deint ? t.nx : t.ny;
    d.srow = deint ? t.sy[0] : t.sx[1];
    const int fs = deint ? 1 : 0, rs = deint ? 0 : 1;  // slots of the flat side and of the rows side
    constexpr int VE = 16 / S;
    bool vec = d.n % VE == 0 && d.srow % VE == 0 && reinterpret_cast<uintptr_t>(pc) % 16 == 0 &&
               reinterpret_cast<uintptr_t>(pa) % 16 == 0;
    d.nbatch = t.nbatch;
    int64_t nb = 1;
    for (int i = 0; i < t.nbatch; ++i) {
        d.bdiv[i] = t.bdiv[i];
        d.bstride_flat[i] = t.bstride[fs][i];
        d.bstride_rows[i] = t.bstride[rs][i];
        vec = vec && t.bstride[0][i] % VE == 0 && t.bstride[1][i] % VE == 0;
        nb *= t.bdiv[i].d;
    }
    // tile: R long positions, a multiple of 32 units (unit = VE positions = k * 16 bytes), payload <= 20 KB so that all
    // of a tile's loads are in flight at once and 8 CTAs fit an SM (profiles/r02_tile_short.md: larger tiles lose 15-20 %)
    static const uint64_t tile_bytes = [] { const char *e = getenv("RC_SHORT_TILE_BYTES"); return e ? (uint64_t)atoll(e) : (uint64_t)SHORT_TILE_BYTES; }();
    const uint32_t g = 32 * VE;
    uint32_t R = (uint32_t)(tile_bytes / ((uint64_t)d.k * S) / g) * g;
    R = std::max<uint32_t>(g, std::min<uint32_t>(R, (d.n + g - 1) / g * g));
    d.R = R;
    d.div_k = FastDiv(d.k);
    d.div_ch = FastDiv(R / (32 * (vec ? VE : 1)));
    d.tiles = (d.n + R - 1) / R;
    d.div_tiles = FastDiv(d.tiles);
    const int64_t total = (int64_t)d.tiles * nb;
    if (total >= (1ll << 31)) return false;
    d.total = (uint32_t)total;
    const size_t smem = (size_t)(R / VE) * (16 * d.k + (S == 8 ? 8 : 4)) + 16;
    unsigned char *flat = deint ? static_cast<unsigned char *>(const_cast<void *>(pa)) : static_cast<unsigned char *>(pc);
    unsigned char *rows = deint ? static_cast<unsigned char *>(pc) : static_cast<unsigned char *>(const_cast<void *>(pa));
    auto go = [&](auto kern) {
        if (smem > 48 * 1024) RC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<d.total, SHORT_THREADS, smem, dev->stream>>>(d, flat, rows);
    };
    if (vec) { if (deint) go(ew_tile_short_kernel<S, true, true>); else go(ew_tile_short_kernel<S, true, false>); }
    else { if (deint) go(ew_tile_short_kernel<S, false, true>); else go(ew_tile_short_kernel<S, false, false>); }
    after_launch(dev, "ew_tile_short_kernel");
    return true;
}

// RC_TILE_WIDE=0: 8-byte permuted copies stay on the square 64 x 64 tile (default: the 128 x 32 tile of rc_tile_wide.cuh)
inline bool tile_wide_enabled() {
    static bool v = [] { const char *e = getenv("RC_TILE_WIDE"); return !(e && e[0] == '0'); }();
    return v;
}

template <class T, int TX, int TY, int NT, class TD>
bool launch_tile_wide(rc_device *dev, const TD &t, T *pc, const T *pa) {
    WideDesc w;
    std::memset(&w, 0, sizeof(w));
    w.nx = t.nx;
    w.ny = t.ny;
    w.tiles_x = (t.nx + TX - 1) / TX;
    w.tiles_y = (t.ny + TY - 1) / TY;
    w.div_tx = FastDiv(w.tiles_x);
    w.div_ty = FastDiv(w.tiles_y);
    w.nbatch = t.nbatch;
    int64_t nb = 1;
    for (int i = 0; i < t.nbatch; ++i) {
        w.bdiv[i] = t.bdiv[i];
        w.bstride_c[i] = t.bstride[0][i];
        w.bstride_a[i] = t.bstride[1][i];
        nb *= t.bdiv[i].d;
    }
    w.sx_a = t.sx[1];
    w.sy_c = t.sy[0];
    const int64_t tiles = (int64_t)w.tiles_x * w.tiles_y * nb;
    if (tiles >= (1ll << 31)) return false;
    w.total_tiles = (uint32_t)tiles;
    const size_t smem = sizeof(T) * TX * (TY + 1);
    if (smem > 48 * 1024)
        RC_CUDA(cudaFuncSetAttribute(ew_tile_wide_kernel<T, TX, TY, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ew_tile_wide_kernel<T, TX, TY, NT><<<w.total_tiles, NT, smem, dev->stream>>>(w, pc, pa);
    after_launch(dev, "ew_tile_wide_kernel");
    return true;
}

// RC_EW_OUTER=0 keeps write-only outer ops on the flat kernel (experiments)
inline bool outer_enabled() {
    static bool v = [] { const char *e = getenv("RC_EW_OUTER"); return !(e && e[0] == '0'); }();
    return v;
}
// RC_TILE_NARROW=0 sends 1- / 2-byte permuted copies back to the one-element-per-lane tile kernels (experiments)
inline bool tile_narrow_enabled() {
    static bool v = [] { const char *e = getenv("RC_TILE_NARROW"); return !(e && e[0] == '0'); }();
    return v;
}

// default choice between the two tile kernels (filled in from the measurement: profiles/r02_tile_bulk.md)
inline bool tile_bulk_default(uint32_t total_tiles, int sm_count) { return false && total_tiles >= 8u * (uint32_t)sm_count; }

// RC_TILE_BULK = 0 / 1 forces the TMA copy kernel off / on where it is eligible (default: see tile_bulk_default)
inline int tile_bulk_mode() {
    static int v = [] { const char *e = getenv("RC_TILE_BULK"); return e ? atoi(e) : -1; }();
    return v;
}

// ---------------------------------------------------------------------------------------------
// rectangular tile kernel: one of X / Y is short and the other long.  The square kernel above
// would leave most of its 64 x 64 slots predicated off (nx = 17: 27 % of the lanes carry data), so here the tile is
// wx x wy with the short extent taken whole and wx * wy <= 4096, and each role walks the tile by a LINEAR index:
//   staged operand   i -> (x, y) = divmod(i, wy): consecutive lanes run along Y, its contiguous axis, and wrap
//                    to the next x, so a short Y still gives full warps;
//   direct / output  i -> (y, x) = divmod(i, wx): consecutive lanes run along X and wrap to the next y (for an
//                    output that is contiguous over (y, x) the whole tile is one contiguous run).
// Same load discipline as the square kernel: 16 slots per thread and operand, all predicated LDGs first.
// ---------------------------------------------------------------------------------------------
constexpr int RECT_ELEMS = TILE_X * TILE_Y;
constexpr int RECT_SLOTS = RECT_ELEMS / (TILE_WARPS * 32);
constexpr int RECT_LONG = 128;   // smallest long extent
// Which short extents take this kernel -- measured against the square tile and the flat kernel on (k, n) <-> (n, k)
// copies (scripts/probe_smalldim.py; profiles/r01_results/probe_smalldim_rect*.txt).  The linear walk costs a division
// per slot and phase: ~40 instructions per element, so 8-byte elements reach the DRAM bound (6.1-6.4 TB/s for every
// k <= 32) and 4-byte elements are issue-bound at 4.1-4.35 TB/s.
//   short Y (the staged operand's contiguous runs are short), runs of 16 bytes .. 32 (f64) / 31 (f32) elements:
//     f64 k = 3 / 8 / 17 / 24: 3.5 / 2.5 / 4.7 / 5.6 -> 6.4 / 6.4 / 6.2 / 6.1 TB/s; f32 k = 8 / 16 / 24: 1.6 / 2.6 / 3.7 ->
//     4.4 / 4.3 / 4.1 TB/s; from k = 32 (f32) / 48 (f64) the square tile is ahead;
//   short X (short output rows): f64 16 <= k <= 31 (k = 16 / 24: 4.5 / 5.5 -> 6.4 / 6.2 TB/s; below 16 the flat kernel
//     writes the rows at 6.2-6.4 TB/s), f32 3 <= k <= 31 (flat 3.75 -> 4.1-4.25 TB/s).
// 1- and 2-byte elements keep their previous paths (not measured).
inline int rect_knob(const char *name, int dflt) {  // experiments only
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
inline bool rect_short_y(uint32_t nx, uint32_t ny, size_t esz) {
    static const int max8 = rect_knob("RC_RECT_MAX_Y8", 32), max4 = rect_knob("RC_RECT_MAX_Y4", 31);
    return tile_rect_enabled() && (esz == 8 || esz == 4) && ny <= (uint32_t)(esz == 8 ? max8 : max4) &&
           (size_t)ny * esz >= 16 && nx >= (uint32_t)RECT_LONG;
}
inline bool rect_short_x(uint32_t nx, uint32_t ny, size_t esz) {
    static const int lo8 = rect_knob("RC_RECT_X_LO8", 16), hi8 = rect_knob("RC_RECT_X_HI8", 31);
    static const int lo4 = rect_knob("RC_RECT_X_LO4", 3), hi4 = rect_knob("RC_RECT_X_HI4", 31);
    const uint32_t lo = esz == 8 ? lo8 : lo4, hi = esz == 8 ? hi8 : hi4;
    return tile_rect_enabled() && (esz == 8 || esz == 4) && nx >= lo && nx <= hi && ny >= (uint32_t)RECT_LONG;
}

struct TileRectDesc : TileDesc {
    uint32_t wx, wy;         // tile extents along X and Y (wx * wy <= RECT_ELEMS)
    uint32_t pitch;          // shared-memory row pitch (elements) of a staged tile: s[x * pitch + y]
    FastDiv div_wx, div_wy;
    int32_t sx32[3], sy32[3];  // sx / sy again: offsets inside one tile fit 32 bits (checked by the launcher)
};

__device__ __forceinline__ uint32_t rect_tid() {
    uint32_t t = threadIdx.x;
    asm volatile("" : "+r"(t));  // opaque to common-subexpression elimination
    return t;
}

// all loads of one operand of the rectangular tile: staged role (x, y) = divmod(i, wy) with unit stride along Y,
// direct role (y, x) = divmod(i, wx); a constant operand keeps `kv`
template <class T>
__device__ __forceinline__ void rect_load(Pack<T, 1> (&r)[RECT_SLOTS], const T *p, T kv, bool staged, bool direct,
                                          const TileRectDesc &d, int k, uint32_t ex, uint32_t ey) {
    const uint32_t tid = rect_tid();
#pragma unroll
    for (int s = 0; s < RECT_SLOTS; ++s) r[s].v[0] = kv;
    if (staged) {
        const int32_t sx = d.sx32[k];
#pragma unroll
        for (int s = 0; s < RECT_SLOTS; ++s) {
            uint32_t x, y;
            d.div_wy.divmod(tid + s * (TILE_WARPS * 32), x, y);
            ld_stream_pred<T, 1>(r[s], p + ((int32_t)x * sx + (int32_t)y), x < ex && y < ey);
        }
    } else if (direct) {
        const int32_t sx = d.sx32[k], sy = d.sy32[k];
#pragma unroll
        for (int s = 0; s < RECT_SLOTS; ++s) {
            uint32_t x, y;
            d.div_wx.divmod(tid + s * (TILE_WARPS * 32), y, x);
            ld_stream_pred<T, 1>(r[s], p + ((int32_t)y * sy + (int32_t)x * sx), x < ex && y < ey);
        }
    }
}

// shared-memory accessors by 32-bit shared-window address (the compiler re-derived the window base and re-read the
// pitch from the parameters for every slot when handed a pointer)
template <class T>
__device__ __forceinline__ void rect_sts(uint32_t addr, const T &v) {
    if constexpr (sizeof(T) == 8) {
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(*reinterpret_cast<const unsigned long long *>(&v)) : "memory");
    } else if constexpr (sizeof(T) == 4) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(*reinterpret_cast<const unsigned int *>(&v)) : "memory");
    } else if constexpr (sizeof(T) == 2) {
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short *>(&v)) : "memory");
    } else {
        static_assert(sizeof(T) == 1, "element size");
        const unsigned short t = *reinterpret_cast<const unsigned char *>(&v);
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "h"(t) : "memory");
    }
}
template <class T>
__device__ __forceinline__ T rect_lds(uint32_t addr) {
    T v;
    if constexpr (sizeof(T) == 8) {
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(*reinterpret_cast<unsigned long long *>(&v)) : "r"(addr) : "memory");
    } else if constexpr (sizeof(T) == 4) {
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(*reinterpret_cast<unsigned int *>(&v)) : "r"(addr) : "memory");
    } else if constexpr (sizeof(T) == 2) {
        asm volatile("ld.shared.b16 %0, [%1];" : "=h"(*reinterpret_cast<unsigned short *>(&v)) : "r"(addr) : "memory");
    } else {
        static_assert(sizeof(T) == 1, "element size");
        unsigned short t;
        asm volatile("ld.shared.u8 %0, [%1];" : "=h"(t) : "r"(addr) : "memory");
        *reinterpret_cast<unsigned char *>(&v) = (unsigned char)t;
    }
    return v;
}
__device__ __forceinline__ uint32_t rect_keep(uint32_t v) {
    asm volatile("" : "+r"(v));  // pinned in a register
    return v;
}

// registers of a staged operand -> shared memory s[x * pitch + y], same walk as its loads
template <class T>
__device__ __forceinline__ void rect_stage(uint32_t sm, const Pack<T, 1> (&r)[RECT_SLOTS], const TileRectDesc &d,
                                           uint32_t ex, uint32_t ey) {
    const uint32_t tid = rect_tid();
    const uint32_t pitch = rect_keep(d.pitch);
#pragma unroll
    for (int s = 0; s < RECT_SLOTS; ++s) {
        uint32_t x, y;
        d.div_wy.divmod(tid + s * (TILE_WARPS * 32), x, y);
        if (x < ex && y < ey)  // also keeps the slot inside the wx x pitch array
            rect_sts<T>(sm + (x * pitch + y) * (uint32_t)sizeof(T), r[s].v[0]);
    }
}

template <class F>
__global__ void __launch_bounds__(TILE_WARPS * 32) ew_tile_rect_kernel(const __grid_constant__ TileRectDesc d,
                                                                        typename F::TO *c, const typename F::TA *a,
                                                                        const typename F::TB *b, int mode_a, int mode_b,
                                                                        EwConst<typename F::TA> ka,
                                                                        EwConst<typename F::TB> kb, ew_params_t<F> prm) {
    using TA = typename F::TA;
    using TB = typename F::TB;
    using TO = typename F::TO;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // staged tiles by shared-window address: a at the start, b behind it (16-byte aligned)
    const uint32_t sa = rect_keep((uint32_t)__cvta_generic_to_shared(smem_raw));
    const uint32_t sb = rect_keep(
        sa + ((mode_a == TILE_STAGED) ? (((uint32_t)sizeof(TA) * d.wx * d.pitch + 15u) & ~15u) : 0u));

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
    const uint32_t x0 = tx * d.wx, y0 = ty * d.wy;
    const uint32_t ex = min(d.wx, d.nx - x0), ey = min(d.wy, d.ny - y0);  // valid extent of this tile

    const bool stg_a = (F::NIN >= 1) && mode_a == TILE_STAGED, dir_a = (F::NIN >= 1) && mode_a == TILE_DIRECT;
    const bool stg_b = (F::NIN >= 2) && mode_b == TILE_STAGED, dir_b = (F::NIN >= 2) && mode_b == TILE_DIRECT;
    const TA *pa = a + base[1] + (int64_t)x0 * d.sx[1] + (int64_t)y0 * d.sy[1];
    const TB *pb = b + base[2] + (int64_t)x0 * d.sx[2] + (int64_t)y0 * d.sy[2];

    // Slot coordinates are recomputed in every phase from a laundered thread index: kept live across the phases
    // (the compiler's choice when it can prove them equal) they cost 64 registers and half the occupancy.
    // Index arithmetic is what bounds this kernel (first version: 68 instructions per element, 32 % of them IMAD,
    // issue-active 65-75 % at 30-55 % DRAM -- profiles/r01_ncu_rect_summary.json), hence 32-bit offsets inside the
    // tile, the known unit strides (staged: along Y, output: along X) and one loop per role.
    Pack<TA, 1> ra[RECT_SLOTS];
    Pack<TB, 1> rb[RECT_SLOTS];
    if constexpr (F::NIN >= 1) rect_load<TA>(ra, pa, ka.v, stg_a, dir_a, d, 1, ex, ey);
    if constexpr (F::NIN >= 2) rect_load<TB>(rb, pb, kb.v, stg_b, dir_b, d, 2, ex, ey);
    if constexpr (F::NIN >= 1)
        if (stg_a) rect_stage<TA>(sa, ra, d, ex, ey);
    if constexpr (F::NIN >= 2)
        if (stg_b) rect_stage<TB>(sb, rb, d, ex, ey);
    __syncthreads();

    // phase 2: (y, x) = divmod(i, wx).  No branch around a slot: an invalid slot reads shared-memory word 0 and
    // skips only its store (branches cost a convergence barrier and a re-derived base pointer per slot).
    TO *pc = c + base[0] + (int64_t)y0 * d.sy[0] + x0;
    asm volatile("" : "+l"(pc));  // materialised once, not re-derived per slot
    const uint32_t tid = rect_tid();
    const int32_t syc = d.sy32[0];
    const uint32_t pitch = rect_keep(d.pitch);
#pragma unroll
    for (int s = 0; s < RECT_SLOTS; ++s) {
        uint32_t yd, xd;
        d.div_wx.divmod(tid + s * (TILE_WARPS * 32), yd, xd);
        const bool ok = yd < ey && xd < ex;
        const uint32_t si = ok ? xd * pitch + yd : 0u;
        const TA va = stg_a ? rect_lds<TA>(sa + si * (uint32_t)sizeof(TA)) : ra[s].v[0];
        TO out;
        if constexpr (F::NIN > 1) {
            const TB vb = stg_b ? rect_lds<TB>(sb + si * (uint32_t)sizeof(TB)) : rb[s].v[0];
            out = ew_apply<F>(va, vb, prm);
        } else {
            out = F::apply(va);
        }
        if (ok) {
            Pack<TO, 1> po;
            po.v[0] = out;
            st_stream<TO, 1>(pc + ((int32_t)yd * syc + (int32_t)xd), po);
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
    const void *params = nullptr;           // host copy of F::Params for functors that declare one
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
    constexpr size_t minsz = sizeof(TO) < sizeof(TA) ? (sizeof(TO) < sizeof(TB) ? sizeof(TO) : sizeof(TB))
                                                     : (sizeof(TA) < sizeof(TB) ? sizeof(TA) : sizeof(TB));
    // elements per pack: 16 bytes of the widest type -- or 32 bytes (256-bit LDG / STG) when the operand types differ in
    // size (casts, comparisons): the narrow side would otherwise move 2-4 bytes per thread and instruction
    // (u8 -> f32 measured 3.7 TB/s with 4-byte loads)
    constexpr int V = ALLOW_VEC ? (int)((maxsz != minsz ? 32 : 16) / maxsz) : 1;

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
    ew_params_t<F> prm;
    if constexpr (!std::is_same<ew_params_t<F>, EwNoParams>::value) {
        RC_CHECK(args.params != nullptr, RC_ERR_INVALID_VALUE, "internal: functor parameters missing");
        std::memcpy(&prm, args.params, sizeof(prm));
    }

    auto stride_of = [&](int s, int i) -> int64_t { return s < 0 ? 0 : c.stride[s][i]; };

    // ---- a long 1-D run whose length is not a multiple of the pack: packs for the body, elements for the tail (a flat f32
    // add of 2^26 - 3 elements ran wholly on the one-element path: 5.9 TB/s against 6.7 with packs) ----
    if constexpr (V > 1) {
        if (c.ndim == 1 && c.shape[0] % V != 0 && c.shape[0] >= 64 * 1024 && c.stride[0][0] == 1) {
            bool unit = true;
            for (int s = 1; s < c.nops; ++s) unit = unit && (c.stride[s][0] == 1 || c.stride[s][0] == 0);
            if (unit) {
                CanonEw head = c, tail = c;
                head.shape[0] = c.shape[0] - c.shape[0] % V;
                tail.shape[0] = c.shape[0] % V;
                for (int s = 0; s < c.nops; ++s) tail.base[s] = c.base[s] + head.shape[0] * c.stride[s][0];
                ew_launch_part<F, ALLOW_TILE, ALLOW_VEC>(dev, head, args);
                ew_launch_part<F, ALLOW_TILE, ALLOW_VEC>(dev, tail, args);
                return;
            }
        }
    }

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
    if (!vec_ok && c.ndim >= 2 && c.stride[0][0] == 1) {
        int ydim = -1;
        // the square tile needs both extents above its thresholds; the rectangular one takes its own shapes
        const bool square_x_ok = c.shape[0] >= tile_min_x(sizeof(TO));
        const uint32_t nx32 = (uint32_t)std::min<int64_t>(c.shape[0], 1u << 30);
        // a one-operand word copy with one short extent whose side is flat: ew_tile_short_kernel (rc_tile_short.cuh)
        auto short_kind = [&](int s, int i) -> int {
            if constexpr (NIN == 1 && is_word_copy<F>::value) {
                if (!tile_short_enabled() || s < 0) return 0;
                const int64_t nx = c.shape[0], ny = c.shape[i], kmax = tile_short_max_k(sizeof(TO));
                if (ny >= 2 && ny <= kmax && nx >= RECT_LONG && nx < (1ll << 31) && c.stride[s][0] == ny) return 1;
                if (nx >= 2 && nx <= kmax && ny >= RECT_LONG && ny < (1ll << 31) && c.stride[0][i] == nx) return 2;
            }
            return 0;
        };
        auto unit_dim = [&](int s) {
            if (s < 0) return -1;
            if (c.stride[s][0] == 0 || c.stride[s][0] == 1) return -1;  // already fine along X
            const size_t esz = s == slot_a ? sizeof(TA) : sizeof(TB);
            for (int i = 1; i < c.ndim; ++i) {
                if (c.stride[s][i] != 1) continue;
                const uint32_t ny32 = (uint32_t)std::min<int64_t>(c.shape[i], 1u << 30);
                if (short_kind(s, i)) return i;
                if ((square_x_ok && c.shape[i] >= tile_min_y(esz)) || rect_short_y(nx32, ny32, esz) ||
                    rect_short_x(nx32, ny32, sizeof(TO)))
                    return i;
            }
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
                // word copy, one short extent on a flat side
                if constexpr (NIN == 1 && is_word_copy<F>::value) {
                    const int kind = tm_a == TILE_STAGED ? short_kind(slot_a, ydim) : 0;
                    if (kind && launch_tile_short<(int)sizeof(TO)>(dev, t, kind, pc, pa)) return;
                }
                // one short and one long extent: rectangular tile, the short extent taken whole
                const size_t esz_staged = (ya == ydim && slot_a >= 0) ? sizeof(TA) : sizeof(TB);
                if constexpr (sizeof(TA) <= 8 && sizeof(TB) <= 8 && sizeof(TO) <= 8)  // 16-byte elements (c64): square tile only
                if (rect_short_y(t.nx, t.ny, esz_staged) || rect_short_x(t.nx, t.ny, sizeof(TO))) {
                    TileRectDesc r;
                    std::memset(&r, 0, sizeof(r));
                    static_cast<TileDesc &>(r) = t;
                    if (t.nx <= t.ny) {
                        r.wx = t.nx;
                        r.wy = std::min<uint32_t>(t.ny, (uint32_t)(RECT_ELEMS / r.wx) & ~31u);
                        r.pitch = r.wy + ((31u - (r.wy & 31u)) & 31u);  // pitch = 31 (mod 32): lanes that wrap to the
                                                                        // next y miss the banks of the first ones
                    } else {
                        r.wy = t.ny;
                        r.wx = std::min<uint32_t>(t.nx, (uint32_t)(RECT_ELEMS / r.wy) & ~31u);
                        r.pitch = r.wy | 1u;
                    }
                    r.tiles_x = (t.nx + r.wx - 1) / r.wx;
                    r.tiles_y = (t.ny + r.wy - 1) / r.wy;
                    r.div_tx = FastDiv(r.tiles_x);
                    r.div_ty = FastDiv(r.tiles_y);
                    r.div_wx = FastDiv(r.wx);
                    r.div_wy = FastDiv(r.wy);
                    int64_t rect_tiles = (int64_t)r.tiles_x * r.tiles_y * nb;
                    bool fits32 = true;  // offsets inside one tile as 32-bit integers
                    for (int k = 0; k < 3; ++k) {
                        const int64_t ax = t.sx[k] < 0 ? -t.sx[k] : t.sx[k], ay = t.sy[k] < 0 ? -t.sy[k] : t.sy[k];
                        fits32 = fits32 && ax < (1ll << 31) && ay < (1ll << 31) &&
                                 (int64_t)(r.wx - 1) * ax + (int64_t)(r.wy - 1) * ay < (1ll << 31);
                        r.sx32[k] = (int32_t)t.sx[k];
                        r.sy32[k] = (int32_t)t.sy[k];
                    }
                    if (rect_tiles < (1ll << 31) && fits32) {
                        r.total_tiles = (uint32_t)rect_tiles;
                        size_t smem = 0;
                        if (tm_a == TILE_STAGED) smem += (sizeof(TA) * r.wx * r.pitch + 15) & ~(size_t)15;
                        if (tm_b == TILE_STAGED) smem += (sizeof(TB) * r.wx * r.pitch + 15) & ~(size_t)15;
                        if (smem > 48 * 1024)
                            RC_CUDA(cudaFuncSetAttribute(ew_tile_rect_kernel<F>,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        ew_tile_rect_kernel<F><<<r.total_tiles, TILE_WARPS * 32, smem, dev->stream>>>(r, pc, pa, pb, tm_a,
                                                                                                    tm_b, ka, kb, prm);
                        after_launch(dev, "ew_tile_rect_kernel");
                        return;
                    }
                }
                // 1- / 2-byte permuted COPY: word-granular tile (rc_tile_narrow.cuh)
                if constexpr (NIN == 1 && (std::is_same<F, FIdentity<uint8_t>>::value || std::is_same<F, FIdentity<uint16_t>>::value)) {
                    constexpr int ESZ = (int)sizeof(TO), E = 4 / ESZ;
                    bool ok = tile_narrow_enabled() && tm_a == TILE_STAGED && t.sx[0] == 1 && t.sy[1] == 1 && t.nx % E == 0 &&
                              t.ny % E == 0 && (int64_t)t.nx * ESZ >= 64 && (int64_t)t.ny * ESZ >= 64 && t.sx[1] % E == 0 &&
                              t.sy[0] % E == 0 && reinterpret_cast<uintptr_t>(pc) % 4 == 0 && reinterpret_cast<uintptr_t>(pa) % 4 == 0;
                    for (int i = 0; i < t.nbatch && ok; ++i) ok = t.bstride[0][i] % E == 0 && t.bstride[1][i] % E == 0;
                    if (ok) {
                        NarrowDesc nd;
                        std::memset(&nd, 0, sizeof(nd));
                        nd.nx = t.nx;
                        nd.ny = t.ny;
                        nd.tiles_x = (t.nx + NW_WORDS_X * E - 1) / (NW_WORDS_X * E);
                        nd.tiles_y = (t.ny + NW_WORDS_Y * E - 1) / (NW_WORDS_Y * E);
                        nd.div_tx = FastDiv(nd.tiles_x);
                        nd.div_ty = FastDiv(nd.tiles_y);
                        nd.nbatch = t.nbatch;
                        int64_t nb2 = 1;
                        for (int i = 0; i < t.nbatch; ++i) {
                            nd.bdiv[i] = t.bdiv[i];
                            nd.bstride_c[i] = t.bstride[0][i];
                            nd.bstride_a[i] = t.bstride[1][i];
                            nb2 *= t.bdiv[i].d;
                        }
                        nd.sx_a = t.sx[1];
                        nd.sy_c = t.sy[0];
                        const int64_t tiles = (int64_t)nd.tiles_x * nd.tiles_y * nb2;
                        if (tiles < (1ll << 31)) {
                            nd.total_tiles = (uint32_t)tiles;
                            ew_tile_narrow_kernel<ESZ><<<nd.total_tiles, NW_WARPS * 32, 0, dev->stream>>>(
                                nd, reinterpret_cast<unsigned char *>(pc), reinterpret_cast<const unsigned char *>(pa));
                            after_launch(dev, "ew_tile_narrow_kernel");
                            return;
                        }
                    }
                }
                // 8-byte permuted COPY, both extents large: wide tile (longer write runs; rc_tile_wide.cuh)
                if constexpr (NIN == 1 && std::is_same<F, FIdentity<uint64_t>>::value) {
                    if (tile_wide_enabled() && tm_a == TILE_STAGED && t.sx[0] == 1 && t.sy[1] == 1 && t.nx >= 256 && t.ny >= 64 &&
                        tile_bulk_mode() != 1 && launch_tile_wide<TO, 128, 32, 256>(dev, t, pc, pa))
                        return;
                }
                // 8-byte permuted COPY of whole 64 x 64 tiles: the TMA kernel (rc_tile_bulk.cuh) where the tensor maps exist
                if constexpr (NIN == 1 && std::is_same<F, FIdentity<uint64_t>>::value) {
                    const int bulk = tile_bulk_mode();
                    if ((bulk == 1 || (bulk < 0 && tile_bulk_default(t.total_tiles, dev->sm_count))) && tm_a == TILE_STAGED &&
                        t.nx % BK_T == 0 && t.ny % BK_T == 0 && t.sx[0] == 1 && t.sy[1] == 1 && t.nbatch <= 3) {
                        CUtensorMap map_src, map_dst;
                        TmaTileDesc td;
                        std::memset(&td, 0, sizeof(td));
                        td.total_tiles = t.total_tiles;
                        td.nbatch = t.nbatch;
                        td.div_ty = t.div_ty;
                        td.div_tx = t.div_tx;
                        std::vector<TmaDim> sd, dd;
                        sd.push_back(TmaDim{t.nx, t.sx[1], (uint32_t)BK_T, 0});
                        dd.push_back(TmaDim{t.ny, t.sy[0], (uint32_t)BK_T, 0});
                        for (int i = 0; i < t.nbatch; ++i) {
                            td.bdiv[i] = t.bdiv[i];
                            sd.push_back(TmaDim{t.bdiv[i].d, t.bstride[1][i], 1u, 1 + i});
                            dd.push_back(TmaDim{t.bdiv[i].d, t.bstride[0][i], 1u, 1 + i});
                        }
                        if (tma_make_map(&map_src, const_cast<TA *>(pa), t.ny, (uint32_t)BK_T, sd, td.src_slot, CU_TENSOR_MAP_SWIZZLE_NONE) &&
                            tma_make_map(&map_dst, pc, t.nx, 16u, dd, td.dst_slot, CU_TENSOR_MAP_SWIZZLE_128B)) {
                            RC_CUDA(cudaFuncSetAttribute(ew_tile_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BK_SMEM));
                            const uint32_t grid = std::min<uint32_t>(t.total_tiles, (uint32_t)dev->sm_count);
                            ew_tile_tma_kernel<0><<<grid, BK_THREADS, BK_SMEM, dev->stream>>>(map_src, map_dst, td);
                            after_launch(dev, "ew_tile_tma_kernel");
                            return;
                        }
                    }
                }
                if (square_x_ok && t.ny >= (uint32_t)tile_min_y(esz_staged)) {
                    size_t smem = 0;
                    if (tm_a == TILE_STAGED) smem += sizeof(TA) * TILE_X * (TILE_Y + 1);
                    if (tm_b == TILE_STAGED) smem += sizeof(TB) * TILE_X * (TILE_Y + 1);
                    if (smem > 48 * 1024)  // opt in to > 48 KB dynamic shared memory (per device)
                        RC_CUDA(cudaFuncSetAttribute(ew_tile_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)smem));
                    ew_tile_kernel<F><<<t.total_tiles, TILE_WARPS * 32, smem, dev->stream>>>(t, pc, pa, pb, tm_a, tm_b, ka, kb, prm);
                    after_launch(dev, "ew_tile_kernel");
                    return;
                }
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
            if (slot_a >= 0 && c.stride[slot_a][0] == 0) mode_a = MODE_SPLAT;
            if (slot_b >= 0 && c.stride[slot_b][0] == 0) mode_b = MODE_SPLAT;
            // write-only outer op: one operand is a scalar per row, the other does not change from row to row
            bool outer_done = false;
            if constexpr (NIN == 2) {
                if (c.ndim == 2 && c.shape[1] < (1ll << 31) && c.shape[1] >= 2 && outer_enabled()) {
                    const bool row_a = mode_a == MODE_SPLAT && c.stride[slot_a][1] != 0;
                    const bool row_b = mode_b == MODE_SPLAT && c.stride[slot_b][1] != 0;
                    const bool kept_a = mode_a == MODE_CONST || (mode_a == MODE_MEM && c.stride[slot_a][1] == 0);
                    const bool kept_b = mode_b == MODE_CONST || (mode_b == MODE_MEM && c.stride[slot_b][1] == 0);
                    if ((row_a && kept_b) || (row_b && kept_a)) {
                        EwRowsDesc rd;
                        std::memset(&rd, 0, sizeof(rd));
                        rd.n0 = (uint32_t)(c.shape[0] / V);
                        rd.n1 = (uint32_t)c.shape[1];
                        for (int k = 0; k < 3; ++k) {
                            rd.s0[k] = stride_of(slots[k], 0) * V;
                            rd.s1[k] = stride_of(slots[k], 1);
                        }
                        rd.rows_per_cta = EW_OUTER_ROWS;
                        const uint32_t chunks = (rd.n0 + EW_BLOCK * EW_UNROLL - 1) / (EW_BLOCK * EW_UNROLL);
                        rd.div_chunks = FastDiv(chunks);
                        const int64_t g2 = ((int64_t)(rd.n1 + rd.rows_per_cta - 1) / rd.rows_per_cta) * chunks;
                        if (g2 < (1ll << 31)) {
                            if (row_a) ew_outer_kernel<F, V, 1><<<(unsigned)g2, EW_BLOCK, 0, dev->stream>>>(rd, pc, pa, pb, mode_b, ka, kb, prm);
                            else ew_outer_kernel<F, V, 2><<<(unsigned)g2, EW_BLOCK, 0, dev->stream>>>(rd, pc, pa, pb, mode_a, ka, kb, prm);
                            after_launch(dev, "ew_outer_kernel");
                            outer_done = true;
                        }
                    }
                }
            }
            if (outer_done) return;
            if (c.ndim == 1) {
                ew_kernel<F, V, 1><<<grid, EW_BLOCK, 0, dev->stream>>>(d, pc, pa, pb, mode_a, mode_b, ka, kb, prm);
            } else if (c.ndim == 2 && c.shape[0] / V >= 512 && c.shape[1] < (1ll << 31) &&
                       // a splat operand that changes from row to row stays on the flat kernel: the rows kernel would be
                       // correct with one row per CTA, but its dependent scalar load at the head of every CTA measured slower
                       // ((n,n)+(n,1) 6.76 -> 6.12 TB/s, outer sum (n,1)+(1,n) 4.38 -> 4.07 TB/s)
                       !(mode_a == MODE_SPLAT && c.stride[slot_a][1] != 0) &&
                       !(mode_b == MODE_SPLAT && c.stride[slot_b][1] != 0)) {
                EwRowsDesc rd;
                std::memset(&rd, 0, sizeof(rd));
                rd.n0 = (uint32_t)(c.shape[0] / V);
                rd.n1 = (uint32_t)c.shape[1];
                bool bcast = false;
                for (int k = 0; k < 3; ++k) {
                    rd.s0[k] = stride_of(slots[k], 0) * V;
                    rd.s1[k] = stride_of(slots[k], 1);
                    if (k > 0 && slots[k] >= 0 && rd.s1[k] == 0) bcast = true;
                }
                // an operand broadcast over the rows is kept in registers for R rows; otherwise one-shot
                // rows per CTA when an operand is broadcast over the rows.  Measured on cfg1 (B200): R = 2 / 4 / 8 / 16 /
                // 32 -> 6.72 / 6.63 / 6.46 / 6.29 / 6.12 TB/s: the re-read of the broadcast operand is served by L2 and
                // costs less than the parallelism lost to longer CTAs, so the default is one row (one-shot grid).
                rd.rows_per_cta = bcast ? ew_rows_per_cta() : 1;
                uint32_t chunks = (rd.n0 + EW_BLOCK * EW_UNROLL - 1) / (EW_BLOCK * EW_UNROLL);
                rd.div_chunks = FastDiv(chunks);
                int64_t groups = (rd.n1 + rd.rows_per_cta - 1) / rd.rows_per_cta;
                int64_t g2 = groups * chunks;
                if (g2 < (1ll << 31)) {
                    ew_rows_kernel<F, V><<<(unsigned)g2, EW_BLOCK, 0, dev->stream>>>(rd, pc, pa, pb, mode_a, mode_b, ka, kb, prm);
                    after_launch(dev, "ew_rows_kernel");
                    return;
                }
                ew_kernel<F, V, 0><<<grid, EW_BLOCK, 0, dev->stream>>>(d, pc, pa, pb, mode_a, mode_b, ka, kb, prm);
            } else {
                // (a compile-time rank-2 instance of the VECTOR kernel was measured too: outer sum unchanged at 4.4 TB/s,
                // (n,n) + (n,1) 6.76 -> 6.51 TB/s -- profiles/r01_results/probe_ew_nd2_all.txt -- so packs keep the
                // runtime-rank walk; the scalar path below is where rank 2 pays)
                ew_kernel<F, V, 0><<<grid, EW_BLOCK, 0, dev->stream>>>(d, pc, pa, pb, mode_a, mode_b, ka, kb, prm);
            }
            after_launch(dev, "ew_kernel");
            return;
        }
    }
    // scalar path: element-rate bound, so 2-D problems take the compile-time rank (f32 a[:, 1:-1] copy 3.8 -> 5.45 TB/s).
    // Elements of <= 4 bytes: 8 instead of 4 items per thread -- 4 loads of 4 bytes per operand do not cover the memory
    // latency (f32 add of operands one element off the pack alignment: 5.8 TB/s against 6.9 for f64)
    constexpr bool NARROW = sizeof(TO) <= 4 && sizeof(TA) <= 4 && sizeof(TB) <= 4;
    constexpr int UN = NARROW ? 2 * EW_UNROLL : EW_UNROLL;
    const uint32_t grid1 = (uint32_t)((items + EW_BLOCK * UN - 1) / (EW_BLOCK * UN));
    if (c.ndim == 2) ew_kernel<F, 1, 2, UN><<<grid1, EW_BLOCK, 0, dev->stream>>>(d, pc, pa, pb, mode_a, mode_b, ka, kb, prm);
    else if (c.ndim == 1) ew_kernel<F, 1, 1, UN><<<grid1, EW_BLOCK, 0, dev->stream>>>(d, pc, pa, pb, mode_a, mode_b, ka, kb, prm);
    else ew_kernel<F, 1, 0, UN><<<grid1, EW_BLOCK, 0, dev->stream>>>(d, pc, pa, pb, mode_a, mode_b, ka, kb, prm);
    after_launch(dev, "ew_kernel");
}

template <class F, bool ALLOW_TILE = true, bool ALLOW_VEC = true>
void ew_launch(rc_device *dev, const CanonEw &c, const EwArgs &args) {
    ew_for_each_part(c, [&](const CanonEw &p) { ew_launch_part<F, ALLOW_TILE, ALLOW_VEC>(dev, p, args); });
}

}  // namespace rc
