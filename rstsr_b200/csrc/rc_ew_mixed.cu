// rc_ew_mixed.cu -- promote_pair FUSED into the kernel for the common mixed operand pairs of + - * /
// (rstsr-core/src/feature_rayon/auto_impl/op_ternary_common.rs:75-139 calls `promote_pair(a, b)` per element and then
// the op on the promoted values; DTypePromoteAPI table: rstsr-dtype-traits/src/promotion.rs:186-300).
//
// The general path of rc_op_mutc_refa_refb_ex casts the operand(s) whose type differs from the compute type K into a
// compact temporary and then runs the one-type kernel: correct for all 121 pairs, but a second pass over memory
// (f32 + f64: 12 + 24 = 36 bytes per element).  For the pairs numerical code actually mixes -- f32 / i32 / i64 against f64
// and i32 against i64, in both orders -- the widening `as` cast happens in registers instead (20 bytes per element), with
// the same kernels (flat packs, rows, tiles): FPromoted<F, K, TA, TB>::apply(a, b) = F<K>::apply((K)a, (K)b).  A C++
// static_cast to a wider arithmetic type IS Rust's `as` for these pairs (value-preserving, or round-to-nearest for
// i64 -> f64), so results are bit-identical to the two-pass path -- tests/test_gpu_mixed.py checks both.
#include "rc_dispatch.cuh"

namespace rc {

template <template <class> class F, class K, class TAin, class TBin>
struct FPromoted {
    using TA = TAin;
    using TB = TBin;
    using TO = typename F<K>::TO;
    static constexpr int NIN = 2;
    RC_FN TO apply(TAin a, TBin b) { return F<K>::apply(static_cast<K>(a), static_cast<K>(b)); }
};

namespace {

template <class K, class TA, class TB>
bool launch_pair(rc_device *dev, rc_binop op, const CanonEw &c, const EwArgs &args) {
    switch (op) {
        case RC_ADD: ew_launch<FPromoted<FAdd, K, TA, TB>>(dev, c, args); return true;
        case RC_SUB: ew_launch<FPromoted<FSub, K, TA, TB>>(dev, c, args); return true;
        case RC_MUL: ew_launch<FPromoted<FMul, K, TA, TB>>(dev, c, args); return true;
        case RC_DIV: ew_launch<FPromoted<FDiv, K, TA, TB>>(dev, c, args); return true;
        default: return false;
    }
}

}  // namespace

// false: this (op, ta, tb) has no fused kernel -- the caller takes the cast-then-op path
bool run_binary_promoted(rc_device *dev, rc_binop op, rc_dtype k, rc_dtype ta, rc_dtype tb, const CanonEw &c,
                         const EwArgs &args) {
    static const bool off = [] { const char *e = getenv("RC_EW_FUSED_PROMOTE"); return e && e[0] == '0'; }();
    if (off) return false;
#define RC_PAIR(KC, KT, AC, AT, BC, BT) \
    if (k == KC && ta == AC && tb == BC) return launch_pair<KT, AT, BT>(dev, op, c, args);
    RC_PAIR(RC_F64, double, RC_F32, float, RC_F64, double)
    RC_PAIR(RC_F64, double, RC_F64, double, RC_F32, float)
    RC_PAIR(RC_F64, double, RC_I32, int32_t, RC_F64, double)
    RC_PAIR(RC_F64, double, RC_F64, double, RC_I32, int32_t)
    RC_PAIR(RC_F64, double, RC_I64, int64_t, RC_F64, double)
    RC_PAIR(RC_F64, double, RC_F64, double, RC_I64, int64_t)
    RC_PAIR(RC_I64, int64_t, RC_I32, int32_t, RC_I64, int64_t)
    RC_PAIR(RC_I64, int64_t, RC_I64, int64_t, RC_I32, int32_t)
#undef RC_PAIR
    return false;
}

}  // namespace rc
