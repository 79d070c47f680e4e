/*
 * rstsr_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the CPU loops the reference's DeviceFaer / DeviceCpuSerial run for the
 * DeviceCuda hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl
 * reference` legs load this library; librstsr_cuda.so never does.
 *
 * Layout algebra (which layouts to iterate) lives in oracle/layout.py; this file holds the data loops
 * that take already-translated layouts:
 *   - IterLayoutColMajor odometer              rstsr-common/src/layout/iterator.rs:23-197
 *   - elementwise closures                     rstsr-core/src/feature_rayon/auto_impl/op_{ternary,binary}_{arithmetic,common}.rs
 *   - loops (serial)                           rstsr-native-impl/src/cpu_serial/{op_with_func,assignment,reduction}.rs
 *   - loops (rayon regimes, OpenMP here)       rstsr-native-impl/src/cpu_rayon/{op_with_func,assignment,reduction}.rs
 *   - unrolled_reduce                          rstsr-native-impl/src/cpu_serial/reduction.rs:44-83
 *   - min/max/cast semantics                   rstsr-dtype-traits/src/{ext_real.rs:32-87, promotion.rs}
 *
 * The reference itself cannot be compiled here (Rust; no rustc/cargo in the image), so parity is pinned
 * on the reference's own known-answer tests (tests/test_oracle_golden.py) -- see DESIGN.md.
 *
 * Build: gcc -O3 -march=x86-64-v2 -fopenmp -shared -fPIC (oracle/Makefile; bench.py rebuilds a -march=native copy on the
 * box it times on).  The `_par` entry points follow the
 * rayon regimes with OpenMP threads and are what bench.py times as the CPU baseline (kind "port").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXD 16

typedef struct {
    int32_t ndim;
    int64_t shape[ORC_MAXD];
    int64_t stride[ORC_MAXD];
    int64_t offset;
} orc_layout;

/* op codes: identical to include/rstsr_cuda.h */
enum { B_ADD = 0, B_SUB, B_MUL, B_DIV, B_REM, B_BITOR, B_BITAND, B_BITXOR, B_SHL, B_SHR, B_MAXIMUM, B_MINIMUM,
       B_FLOOR_DIVIDE, B_POW, B_ATAN2, B_COPYSIGN, B_HYPOT, B_LOGADDEXP, B_NEXTAFTER,
       B_EQ = 32, B_NE, B_LT, B_LE, B_GT, B_GE };
enum { U_NEG = 0, U_NOT, U_ABS, U_SQUARE, U_SIGN, U_SQRT, U_EXP, U_EXPM1, U_LOG, U_LOG2, U_LOG10, U_SIN, U_COS, U_TAN,
       U_ASIN, U_ACOS, U_ATAN, U_SINH, U_COSH, U_TANH, U_ASINH, U_ACOSH, U_ATANH, U_FLOOR, U_CEIL, U_ROUND, U_TRUNC,
       U_RECIPROCAL, U_CONJ, U_REAL, U_IMAG, U_ISNAN = 48, U_ISINF, U_ISFINITE, U_SIGNBIT };
enum { R_SUM = 0, R_PROD, R_MAX, R_MIN, R_MEAN };
enum { T_BOOL = 0, T_I8, T_I16, T_I32, T_I64, T_U8, T_U16, T_U32, T_U64, T_F32, T_F64 };

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* The timed CPU baseline mirrors DeviceFaer::default() = the rayon global pool = every core the process may run on
 * (rstsr-core/src/feature_rayon/device.rs:53-75).  Launchers such as torchrun export OMP_NUM_THREADS=1, which would
 * silently turn the baseline into a single-thread run: the bench sets the count explicitly. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- IterLayoutColMajor (iterator.rs:122-197): axis 0 fastest; position k -> offset ---- */
typedef struct {
    const orc_layout *l;
    int64_t idx[ORC_MAXD];
    int64_t off;
    int64_t remaining;
} orc_iter;

static int64_t layout_size(const orc_layout *l) {
    int64_t n = 1;
    for (int i = 0; i < l->ndim; ++i) n *= l->shape[i];
    return n;
}

static void iter_init_at(orc_iter *it, const orc_layout *l, int64_t pos) {
    it->l = l;
    it->off = l->offset;
    it->remaining = layout_size(l) - pos;
    for (int i = 0; i < l->ndim; ++i) {
        int64_t d = l->shape[i] ? l->shape[i] : 1;
        it->idx[i] = pos % d;
        pos /= d;
        it->off += it->idx[i] * l->stride[i];
    }
}

static inline void iter_next(orc_iter *it) {
    const orc_layout *l = it->l;
    it->remaining--;
    for (int k = 0; k < l->ndim; ++k) {
        it->idx[k]++;
        it->off += l->stride[k];
        if (it->idx[k] < l->shape[k]) return;
        it->off -= l->stride[k] * l->shape[k];
        it->idx[k] = 0;
    }
}

/* =====================================================================================================
 * Typed section, instantiated once per element type through oracle_typed.inc-style macros below.
 * ===================================================================================================== */
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)

/* ---- integer instantiation ---- */
#define DEFINE_INT(T, UT, SUF, BITS, TMIN, TMAX)                                                                   \
    static inline T CAT(bin, SUF)(int op, T a, T b) {                                                               \
        switch (op) {                                                                                               \
            case B_ADD: return (T)((UT)a + (UT)b); /* wrapping: the reference's CI runs --release */              \
            case B_SUB: return (T)((UT)a - (UT)b);                                                                  \
            case B_MUL: return (T)((uint64_t)(UT)a * (uint64_t)(UT)b);                                              \
            case B_DIV: return b == 0 ? 0 : ((TMIN != 0 && b == (T)-1) ? (T)(0 - (UT)a) : (T)(a / b));              \
            case B_REM: return b == 0 ? 0 : ((TMIN != 0 && b == (T)-1) ? 0 : (T)(a % b));                           \
            case B_BITOR: return (T)(a | b);                                                                        \
            case B_BITAND: return (T)(a & b);                                                                       \
            case B_BITXOR: return (T)(a ^ b);                                                                       \
            case B_SHL: return (T)((UT)a << ((unsigned)b & (BITS - 1)));                                            \
            case B_SHR: return (T)(a >> ((unsigned)b & (BITS - 1)));                                                \
            case B_MAXIMUM: return a < b ? b : a;                                                                   \
            case B_MINIMUM: return b < a ? b : a;                                                                   \
            case B_FLOOR_DIVIDE: {                                                                                  \
                if (b == 0) return 0;                                                                               \
                if (TMIN != 0 && b == (T)-1) return (T)(0 - (UT)a);                                                 \
                T q = (T)(a / b), r = (T)(a % b);                                                                   \
                return (r != 0 && ((r < 0) != (b < 0))) ? (T)(q - 1) : q;                                           \
            }                                                                                                       \
        }                                                                                                           \
        return 0;                                                                                                   \
    }                                                                                                               \
    static inline T CAT(una, SUF)(int op, T a) {                                                                    \
        switch (op) {                                                                                               \
            case U_NEG: return (T)(0 - (UT)a);                                                                      \
            case U_NOT: return (T)~a;                                                                               \
            case U_ABS: return (TMIN != 0 && a < 0) ? (T)(0 - (UT)a) : a;                                           \
            case U_SQUARE: return (T)((uint64_t)(UT)a * (uint64_t)(UT)a);                                           \
            case U_SIGN: return (TMIN != 0) ? (T)(a > 0 ? 1 : (a < 0 ? -1 : 0)) : (T)(a == 0 ? 0 : 1);              \
            case U_CONJ: case U_REAL: return a;                                                                     \
            case U_IMAG: return 0;                                                                                  \
        }                                                                                                           \
        return 0;                                                                                                   \
    }                                                                                                               \
    static inline T CAT(rinit, SUF)(int op) {                                                                       \
        return op == R_PROD ? (T)1 : (op == R_MAX ? (T)TMIN : (op == R_MIN ? (T)TMAX : (T)0));                      \
    }                                                                                                               \
    static inline T CAT(rf, SUF)(int op, T a, T b) {                                                                \
        switch (op) {                                                                                               \
            case R_PROD: return (T)((uint64_t)(UT)a * (uint64_t)(UT)b);                                             \
            case R_MAX: return a < b ? b : a; /* Ord::max */                                                        \
            case R_MIN: return b < a ? b : a;                                                                       \
            default: return (T)((UT)a + (UT)b);                                                                     \
        }                                                                                                           \
    }                                                                                                               \
    static inline T CAT(rout, SUF)(int op, T a, int64_t n) { (void)op; (void)n; return a; }

/* ---- float instantiation ---- */
#define DEFINE_FLT(T, SUF, FMAXF, FMINF, FMODF, FLOORF, POWF, ATAN2F, COPYSIGNF, HYPOTF, LOGF, EXPF, NEXTF, TLOWEST, \
                   THIGHEST, MATH)                                                                                   \
    static inline T CAT(bin, SUF)(int op, T a, T b) {                                                               \
        switch (op) {                                                                                               \
            case B_ADD: return a + b;                                                                               \
            case B_SUB: return a - b;                                                                               \
            case B_MUL: return a * b;                                                                               \
            case B_DIV: return a / b;                                                                               \
            case B_REM: return FMODF(a, b);                                                                         \
            case B_MAXIMUM: return FMAXF(a, b); /* f64::max: NaN-ignoring */                                       \
            case B_MINIMUM: return FMINF(a, b);                                                                     \
            case B_FLOOR_DIVIDE: return FLOORF(a / b);                                                              \
            case B_POW: return POWF(a, b);                                                                          \
            case B_ATAN2: return ATAN2F(a, b);                                                                      \
            case B_COPYSIGN: return COPYSIGNF(a, b);                                                                \
            case B_HYPOT: return HYPOTF(a, b);                                                                      \
            case B_LOGADDEXP: return LOGF(EXPF(a) + EXPF(b));                                                       \
            case B_NEXTAFTER: return NEXTF(a, b);                                                                   \
        }                                                                                                           \
        return 0;                                                                                                   \
    }                                                                                                               \
    static inline T CAT(una, SUF)(int op, T a) {                                                                    \
        switch (op) {                                                                                               \
            case U_NEG: return -a;                                                                                  \
            case U_ABS: return MATH(fabs)(a);                                                                       \
            case U_SQUARE: return a * a;                                                                            \
            case U_SIGN: return a != a ? a : (a > 0 ? (T)1 : (a < 0 ? (T)-1 : (T)0));                               \
            case U_SQRT: return MATH(sqrt)(a);                                                                      \
            case U_EXP: return MATH(exp)(a);                                                                        \
            case U_EXPM1: return MATH(expm1)(a);                                                                    \
            case U_LOG: return MATH(log)(a);                                                                        \
            case U_LOG2: return MATH(log2)(a);                                                                      \
            case U_LOG10: return MATH(log10)(a);                                                                    \
            case U_SIN: return MATH(sin)(a);                                                                        \
            case U_COS: return MATH(cos)(a);                                                                        \
            case U_TAN: return MATH(tan)(a);                                                                        \
            case U_ASIN: return MATH(asin)(a);                                                                      \
            case U_ACOS: return MATH(acos)(a);                                                                      \
            case U_ATAN: return MATH(atan)(a);                                                                      \
            case U_SINH: return MATH(sinh)(a);                                                                      \
            case U_COSH: return MATH(cosh)(a);                                                                      \
            case U_TANH: return MATH(tanh)(a);                                                                      \
            case U_ASINH: return MATH(asinh)(a);                                                                    \
            case U_ACOSH: return MATH(acosh)(a);                                                                    \
            case U_ATANH: return MATH(atanh)(a);                                                                    \
            case U_FLOOR: return MATH(floor)(a);                                                                    \
            case U_CEIL: return MATH(ceil)(a);                                                                      \
            case U_ROUND: return MATH(round)(a);                                                                    \
            case U_TRUNC: return MATH(trunc)(a);                                                                    \
            case U_RECIPROCAL: return (T)1 / a;                                                                     \
            case U_CONJ: case U_REAL: return a;                                                                     \
            case U_IMAG: return 0;                                                                                  \
        }                                                                                                           \
        return 0;                                                                                                   \
    }                                                                                                               \
    static inline T CAT(rinit, SUF)(int op) {                                                                       \
        return op == R_PROD ? (T)1 : (op == R_MAX ? (T)TLOWEST : (op == R_MIN ? (T)THIGHEST : (T)0));               \
    }                                                                                                               \
    static inline T CAT(rf, SUF)(int op, T a, T b) {                                                                \
        switch (op) {                                                                                               \
            case R_PROD: return a * b;                                                                              \
            case R_MAX: return FMAXF(a, b);                                                                         \
            case R_MIN: return FMINF(a, b);                                                                         \
            default: return a + b;                                                                                  \
        }                                                                                                           \
    }                                                                                                               \
    static inline T CAT(rout, SUF)(int op, T a, int64_t n) { return op == R_MEAN ? a / (T)n : a; }

#define MATHF(fn) fn##f
#define MATHD(fn) fn
DEFINE_INT(int8_t, uint8_t, i8, 8, INT8_MIN, INT8_MAX)
DEFINE_INT(int16_t, uint16_t, i16, 16, INT16_MIN, INT16_MAX)
DEFINE_INT(int32_t, uint32_t, i32, 32, INT32_MIN, INT32_MAX)
DEFINE_INT(int64_t, uint64_t, i64, 64, INT64_MIN, INT64_MAX)
DEFINE_INT(uint8_t, uint8_t, u8, 8, 0, UINT8_MAX)
DEFINE_INT(uint16_t, uint16_t, u16, 16, 0, UINT16_MAX)
DEFINE_INT(uint32_t, uint32_t, u32, 32, 0, UINT32_MAX)
DEFINE_INT(uint64_t, uint64_t, u64, 64, 0, UINT64_MAX)
DEFINE_FLT(float, f32, fmaxf, fminf, fmodf, floorf, powf, atan2f, copysignf, hypotf, logf, expf, nextafterf,
           -3.40282346638528859811704183484516925e+38F, 3.40282346638528859811704183484516925e+38F, MATHF)
DEFINE_FLT(double, f64, fmax, fmin, fmod, floor, pow, atan2, copysign, hypot, log, exp, nextafter,
           -1.79769313486231570814527423731704357e+308, 1.79769313486231570814527423731704357e+308, MATHD)

/* ---- loops common to every numeric type ---- */
#define DEFINE_LOOPS(T, SUF)                                                                                         \
    /* c[idx] = f(a[idx], b[idx]); NULL a / b = scalar operand (op_mutc_refa_numb / numa_refb).               */ \
    /* cpu_serial/op_with_func.rs:10-199: same logical index in every operand; order is irrelevant to values.  */ \
    void CAT(orc_binary, SUF)(int op, void *c_, const orc_layout *lc, const T *a, const orc_layout *la, const T *b, \
                               const orc_layout *lb, const T *sa, const T *sb) {                                     \
        int64_t n = layout_size(lc);                                                                                 \
        if (n == 0) return;                                                                                          \
        orc_layout dummy = *lc;                                                                                      \
        for (int i = 0; i < dummy.ndim; ++i) dummy.stride[i] = 0;                                                    \
        dummy.offset = 0;                                                                                            \
        orc_iter ic, ia, ib;                                                                                         \
        iter_init_at(&ic, lc, 0);                                                                                    \
        iter_init_at(&ia, a ? la : &dummy, 0);                                                                       \
        iter_init_at(&ib, b ? lb : &dummy, 0);                                                                       \
        for (int64_t k = 0; k < n; ++k) {                                                                            \
            T x = a ? a[ia.off] : *sa, y = b ? b[ib.off] : *sb;                                                      \
            if (op >= B_EQ) {                                                                                        \
                uint8_t r = 0;                                                                                       \
                switch (op) {                                                                                        \
                    case B_EQ: r = x == y; break;                                                                    \
                    case B_NE: r = x != y; break;                                                                    \
                    case B_LT: r = x < y; break;                                                                     \
                    case B_LE: r = x <= y; break;                                                                    \
                    case B_GT: r = x > y; break;                                                                     \
                    case B_GE: r = x >= y; break;                                                                    \
                }                                                                                                    \
                ((uint8_t *)c_)[ic.off] = r;                                                                         \
            } else {                                                                                                 \
                ((T *)c_)[ic.off] = CAT(bin, SUF)(op, x, y);                                                         \
            }                                                                                                        \
            iter_next(&ic); iter_next(&ia); iter_next(&ib);                                                          \
        }                                                                                                            \
    }                                                                                                                \
    /* rayon regimes of cpu_rayon/op_with_func.rs:13-80 for c = a + b (the timed CPU baseline):                */ \
    /* layouts are the OUTER layouts after translate_to_col_major_with_contig; size_contig the peeled run.      */ \
    void CAT(orc_add_par, SUF)(T *c, const orc_layout *lc, const T *a, const orc_layout *la, const T *b,             \
                                const orc_layout *lb, int64_t size_contig) {                                         \
        int64_t nouter = layout_size(lc);                                                                            \
        if (size_contig >= 16) {                                                                                     \
            if (size_contig < 4096) {                                                                                \
                /* parallel over outer index, serial inner run (:47-56) */                                           \
                _Pragma("omp parallel") {                                                                            \
                    int nt = 1, tid = 0;                                                                             \
                    nt = orc_num_threads_in(); tid = orc_thread_id();                                                \
                    int64_t lo = nouter * tid / nt, hi = nouter * (tid + 1) / nt;                                    \
                    orc_iter ic, ia, ib;                                                                             \
                    iter_init_at(&ic, lc, lo); iter_init_at(&ia, la, lo); iter_init_at(&ib, lb, lo);                 \
                    for (int64_t k = lo; k < hi; ++k) {                                                              \
                        T *cp = c + ic.off; const T *ap = a + ia.off, *bp = b + ib.off;                              \
                        for (int64_t i = 0; i < size_contig; ++i) cp[i] = ap[i] + bp[i];                             \
                        iter_next(&ic); iter_next(&ia); iter_next(&ib);                                              \
                    }                                                                                                \
                }                                                                                                    \
            } else {                                                                                                 \
                /* outer items in sequence, each run split over the pool (:57-67) */                                 \
                orc_iter ic, ia, ib;                                                                                 \
                iter_init_at(&ic, lc, 0); iter_init_at(&ia, la, 0); iter_init_at(&ib, lb, 0);                        \
                _Pragma("omp parallel") {                                                                            \
                    for (int64_t k = 0; k < nouter; ++k) {                                                           \
                        T *cp = c + ic.off; const T *ap = a + ia.off, *bp = b + ib.off;                              \
                        _Pragma("omp for schedule(static) nowait")                                                   \
                        for (int64_t i = 0; i < size_contig; ++i) cp[i] = ap[i] + bp[i];                             \
                        _Pragma("omp barrier")                                                                       \
                        _Pragma("omp single") { iter_next(&ic); iter_next(&ia); iter_next(&ib); }                    \
                    }                                                                                                \
                }                                                                                                    \
            }                                                                                                        \
        } else {                                                                                                     \
            /* fully strided: per-element odometer on all three operands (:68-79); lc/la/lb are the FULL layouts */ \
            _Pragma("omp parallel") {                                                                                \
                int nt = orc_num_threads_in(), tid = orc_thread_id();                                                \
                int64_t lo = nouter * tid / nt, hi = nouter * (tid + 1) / nt;                                        \
                orc_iter ic, ia, ib;                                                                                 \
                iter_init_at(&ic, lc, lo); iter_init_at(&ia, la, lo); iter_init_at(&ib, lb, lo);                     \
                for (int64_t k = lo; k < hi; ++k) {                                                                  \
                    c[ic.off] = a[ia.off] + b[ib.off];                                                               \
                    iter_next(&ic); iter_next(&ia); iter_next(&ib);                                                  \
                }                                                                                                    \
            }                                                                                                        \
        }                                                                                                            \
    }                                                                                                                \
    void CAT(orc_unary, SUF)(int op, void *c_, const orc_layout *lc, const T *a, const orc_layout *la) {             \
        int64_t n = layout_size(lc);                                                                                 \
        if (n == 0) return;                                                                                          \
        orc_iter ic, ia;                                                                                             \
        iter_init_at(&ic, lc, 0);                                                                                    \
        iter_init_at(&ia, la, 0);                                                                                    \
        for (int64_t k = 0; k < n; ++k) {                                                                            \
            T x = a[ia.off];                                                                                         \
            if (op >= U_ISNAN) ((uint8_t *)c_)[ic.off] = CAT(pred, SUF)(op, x);                                      \
            else ((T *)c_)[ic.off] = CAT(una, SUF)(op, x);                                                           \
            iter_next(&ic); iter_next(&ia);                                                                          \
        }                                                                                                            \
    }                                                                                                                \
    void CAT(orc_fill, SUF)(T *c, const orc_layout *lc, const T *v) {                                                \
        int64_t n = layout_size(lc);                                                                                 \
        orc_iter ic;                                                                                                 \
        iter_init_at(&ic, lc, 0);                                                                                    \
        for (int64_t k = 0; k < n; ++k) { c[ic.off] = *v; iter_next(&ic); }                                          \
    }                                                                                                                \
    /* unrolled_reduce (cpu_serial/reduction.rs:44-83): 8 lanes, fixed combination order, <= 7 tail elements */     \
    static T CAT(unrolled_reduce, SUF)(int op, const T *xs, int64_t n) {                                             \
        T acc = CAT(rinit, SUF)(op);                                                                                 \
        T p0 = acc, p1 = acc, p2 = acc, p3 = acc, p4 = acc, p5 = acc, p6 = acc, p7 = acc;                            \
        while (n >= 8) {                                                                                             \
            p0 = CAT(rf, SUF)(op, p0, xs[0]); p1 = CAT(rf, SUF)(op, p1, xs[1]);                                      \
            p2 = CAT(rf, SUF)(op, p2, xs[2]); p3 = CAT(rf, SUF)(op, p3, xs[3]);                                      \
            p4 = CAT(rf, SUF)(op, p4, xs[4]); p5 = CAT(rf, SUF)(op, p5, xs[5]);                                      \
            p6 = CAT(rf, SUF)(op, p6, xs[6]); p7 = CAT(rf, SUF)(op, p7, xs[7]);                                      \
            xs += 8; n -= 8;                                                                                         \
        }                                                                                                            \
        acc = CAT(rf, SUF)(op, acc, CAT(rf, SUF)(op, p0, p4));                                                       \
        acc = CAT(rf, SUF)(op, acc, CAT(rf, SUF)(op, p1, p5));                                                       \
        acc = CAT(rf, SUF)(op, acc, CAT(rf, SUF)(op, p2, p6));                                                       \
        acc = CAT(rf, SUF)(op, acc, CAT(rf, SUF)(op, p3, p7));                                                       \
        for (int64_t i = 0; i < n && i < 7; ++i) acc = CAT(rf, SUF)(op, acc, xs[i]);                                 \
        return acc;                                                                                                  \
    }                                                                                                                \
    /* reduce_all_cpu_serial (cpu_serial/reduction.rs:144-178).  `l` = layout after K-order translation;        */ \
    /* size_contig >= 32: `l` is the outer layout and each item a run; else `l` is the full layout.              */ \
    void CAT(orc_reduce_all, SUF)(int op, const T *a, const orc_layout *l, int64_t size_contig, int64_t n_total,     \
                                   T *out) {                                                                         \
        T acc = CAT(rinit, SUF)(op);                                                                                 \
        int64_t n = layout_size(l);                                                                                  \
        orc_iter it;                                                                                                 \
        iter_init_at(&it, l, 0);                                                                                     \
        if (size_contig >= 32) {                                                                                     \
            for (int64_t k = 0; k < n; ++k) {                                                                        \
                acc = CAT(rf, SUF)(op, acc, CAT(unrolled_reduce, SUF)(op, a + it.off, size_contig));                 \
                iter_next(&it);                                                                                      \
            }                                                                                                        \
        } else {                                                                                                     \
            for (int64_t k = 0; k < n; ++k) { acc = CAT(rf, SUF)(op, acc, a[it.off]); iter_next(&it); }              \
        }                                                                                                            \
        *out = CAT(rout, SUF)(op, acc, n_total);                                                                     \
    }                                                                                                                \
    /* reduce_all_cpu_rayon (cpu_rayon/reduction.rs:20-106), run >= 1024: 1024-element chunks through           */ \
    /* unrolled_reduce, combined per thread then across threads (rayon's combination tree is scheduler-         */ \
    /* dependent; this is one valid association).                                                                */ \
    void CAT(orc_reduce_all_par, SUF)(int op, const T *a, const orc_layout *l, int64_t size_contig,                  \
                                       int64_t n_total, T *out) {                                                    \
        if (size_contig < 1024) { CAT(orc_reduce_all, SUF)(op, a, l, size_contig, n_total, out); return; }           \
        int64_t nouter = layout_size(l);                                                                             \
        T acc = CAT(rinit, SUF)(op);                                                                                 \
        orc_iter it;                                                                                                 \
        iter_init_at(&it, l, 0);                                                                                     \
        int nt = orc_num_threads();                                                                                  \
        T *part = (T *)malloc(sizeof(T) * (size_t)nt * 16);                                                          \
        for (int64_t k = 0; k < nouter; ++k) {                                                                       \
            const T *run = a + it.off;                                                                               \
            int64_t nchunk = (size_contig + 1023) / 1024;                                                            \
            for (int t = 0; t < nt; ++t) part[t * 16] = CAT(rinit, SUF)(op);                                         \
            _Pragma("omp parallel") {                                                                                \
                int tid = orc_thread_id();                                                                           \
                T local = CAT(rinit, SUF)(op);                                                                       \
                _Pragma("omp for schedule(static)")                                                                  \
                for (int64_t ch = 0; ch < nchunk; ++ch) {                                                            \
                    int64_t s = ch * 1024, len = size_contig - s < 1024 ? size_contig - s : 1024;                    \
                    local = CAT(rf, SUF)(op, local, CAT(unrolled_reduce, SUF)(op, run + s, len));                    \
                }                                                                                                    \
                part[tid * 16] = local;                                                                              \
            }                                                                                                        \
            T res = CAT(rinit, SUF)(op);                                                                             \
            for (int t = 0; t < nt; ++t) res = CAT(rf, SUF)(op, res, part[t * 16]);                                  \
            acc = CAT(rf, SUF)(op, acc, res);                                                                        \
            iter_next(&it);                                                                                          \
        }                                                                                                            \
        free(part);                                                                                                  \
        *out = CAT(rout, SUF)(op, acc, n_total);                                                                     \
    }                                                                                                                \
    /* reduce_axes regime (a): reduced part has a contiguous run (cpu_serial/reduction.rs:231-256,             */ \
    /* cpu_rayon/reduction.rs:168-198).  rayon_like != 0 reproduces DeviceFaer's `init + (init + unrolled)`      */ \
    /* association of the single-inner-item case exactly; values are otherwise identical.                        */ \
    void CAT(orc_reduce_axes_a, SUF)(int op, const T *a, T *out, const orc_layout *l_mcd, const orc_layout *l_ocd,   \
                                      const orc_layout *l_sd, int64_t size_sc, int64_t size_s0, int64_t offset,      \
                                      int64_t n_mean, int parallel) {                                                \
        int64_t nm = layout_size(l_mcd), ns = layout_size(l_sd);                                                     \
        _Pragma("omp parallel if (parallel)") {                                                                      \
            int nt = parallel ? orc_num_threads_in() : 1, tid = parallel ? orc_thread_id() : 0;                      \
            int64_t lo = nm * tid / nt, hi = nm * (tid + 1) / nt;                                                    \
            orc_iter im, io;                                                                                         \
            iter_init_at(&im, l_mcd, lo);                                                                            \
            iter_init_at(&io, l_ocd, lo);                                                                            \
            for (int64_t k = lo; k < hi; ++k) {                                                                      \
                T acc = CAT(rinit, SUF)(op);                                                                         \
                orc_iter is;                                                                                         \
                iter_init_at(&is, l_sd, 0);                                                                          \
                for (int64_t j = 0; j < ns; ++j) {                                                                   \
                    int64_t idx = im.off + is.off - offset;                                                          \
                    acc = CAT(rf, SUF)(op, acc, CAT(unrolled_reduce, SUF)(op, a + idx, size_sc));                    \
                    iter_next(&is);                                                                                  \
                }                                                                                                    \
                if (parallel) acc = CAT(rf, SUF)(op, CAT(rinit, SUF)(op), acc); /* rayon .reduce(init, f_sum) */     \
                T before = acc;                                                                                      \
                for (int64_t r = 1; r < size_s0; ++r) acc = CAT(rf, SUF)(op, acc, before);                           \
                out[io.off] = CAT(rout, SUF)(op, acc, n_mean);                                                       \
                iter_next(&im); iter_next(&io);                                                                      \
            }                                                                                                        \
        }                                                                                                            \
    }                                                                                                                \
    /* regime (b): kept part contiguous (cpu_serial/reduction.rs:257-304 chunk 48, cpu_rayon :199-248 chunk 64) */  \
    void CAT(orc_reduce_axes_b, SUF)(int op, const T *a, T *out, const orc_layout *l_md, const orc_layout *l_od,     \
                                      const orc_layout *l_scd, int64_t size_mc, int64_t size_s0, int64_t offset,     \
                                      int64_t n_mean, int parallel) {                                                \
        int64_t nm = layout_size(l_md), ns = layout_size(l_scd);                                                     \
        const int64_t CH = parallel ? 64 : 48;                                                                       \
        T *vacc = (T *)malloc(sizeof(T) * (size_t)(size_mc > 0 ? size_mc : 1));                                      \
        orc_iter im, io;                                                                                             \
        iter_init_at(&im, l_md, 0);                                                                                  \
        iter_init_at(&io, l_od, 0);                                                                                  \
        for (int64_t k = 0; k < nm; ++k) {                                                                           \
            int64_t nchunk = (size_mc + CH - 1) / CH;                                                                \
            _Pragma("omp parallel for schedule(dynamic) if (parallel)")                                              \
            for (int64_t ch = 0; ch < nchunk; ++ch) {                                                                \
                int64_t start = ch * CH, len = size_mc - start < CH ? size_mc - start : CH;                          \
                T *v = vacc + start;                                                                                 \
                for (int64_t i = 0; i < len; ++i) v[i] = CAT(rinit, SUF)(op);                                        \
                orc_iter is;                                                                                         \
                iter_init_at(&is, l_scd, 0);                                                                         \
                for (int64_t j = 0; j < ns; ++j) {                                                                   \
                    const T *row = a + (im.off + is.off - offset) + start;                                           \
                    for (int64_t i = 0; i < len; ++i) v[i] = CAT(rf, SUF)(op, v[i], row[i]);                         \
                    iter_next(&is);                                                                                  \
                }                                                                                                    \
            }                                                                                                        \
            for (int64_t i = 0; i < size_mc; ++i) {                                                                  \
                T acc = vacc[i];                                                                                     \
                for (int64_t r = 1; r < size_s0; ++r) acc = CAT(rf, SUF)(op, acc, vacc[i]);                          \
                out[io.off + i] = CAT(rout, SUF)(op, acc, n_mean);                                                   \
            }                                                                                                        \
            iter_next(&im); iter_next(&io);                                                                          \
        }                                                                                                            \
        free(vacc);                                                                                                  \
    }                                                                                                                \
    /* regime (c): nothing contiguous (cpu_serial/reduction.rs:305-328) */                                           \
    void CAT(orc_reduce_axes_c, SUF)(int op, const T *a, T *out, const orc_layout *l_md, const orc_layout *l_od,     \
                                      const orc_layout *l_sd, int64_t size_s0, int64_t offset, int64_t n_mean) {     \
        int64_t nm = layout_size(l_md), ns = layout_size(l_sd);                                                      \
        orc_iter im, io;                                                                                             \
        iter_init_at(&im, l_md, 0);                                                                                  \
        iter_init_at(&io, l_od, 0);                                                                                  \
        for (int64_t k = 0; k < nm; ++k) {                                                                           \
            T acc = CAT(rinit, SUF)(op);                                                                             \
            orc_iter is;                                                                                             \
            iter_init_at(&is, l_sd, 0);                                                                              \
            for (int64_t j = 0; j < ns; ++j) { acc = CAT(rf, SUF)(op, acc, a[im.off + is.off - offset]); iter_next(&is); } \
            T before = acc;                                                                                          \
            for (int64_t r = 1; r < size_s0; ++r) acc = CAT(rf, SUF)(op, acc, before);                               \
            out[io.off] = CAT(rout, SUF)(op, acc, n_mean);                                                           \
            iter_next(&im); iter_next(&io);                                                                          \
        }                                                                                                            \
    }

static inline int orc_thread_id(void) {
#ifdef _OPENMP
    return omp_get_thread_num();
#else
    return 0;
#endif
}
static inline int orc_num_threads_in(void) {
#ifdef _OPENMP
    return omp_get_num_threads();
#else
    return 1;
#endif
}

#define DEFINE_PRED_INT(T, SUF) static inline uint8_t CAT(pred, SUF)(int op, T a) { (void)op; (void)a; return 0; }
/* OpSignBitAPI writes b.is_positive() (auto_impl/op_binary_common.rs:104): sign bit CLEAR */
#define DEFINE_PRED_FLT(T, SUF)                                             \
    static inline uint8_t CAT(pred, SUF)(int op, T a) {                     \
        switch (op) {                                                       \
            case U_ISNAN: return a != a;                                    \
            case U_ISINF: return isinf(a) != 0;                             \
            case U_ISFINITE: return isfinite(a) != 0;                       \
            case U_SIGNBIT: return !signbit(a);                             \
        }                                                                   \
        return 0;                                                           \
    }
DEFINE_PRED_INT(int8_t, i8) DEFINE_PRED_INT(int16_t, i16) DEFINE_PRED_INT(int32_t, i32) DEFINE_PRED_INT(int64_t, i64)
DEFINE_PRED_INT(uint8_t, u8) DEFINE_PRED_INT(uint16_t, u16) DEFINE_PRED_INT(uint32_t, u32) DEFINE_PRED_INT(uint64_t, u64)
DEFINE_PRED_FLT(float, f32) DEFINE_PRED_FLT(double, f64)

DEFINE_LOOPS(int8_t, i8)
DEFINE_LOOPS(int16_t, i16)
DEFINE_LOOPS(int32_t, i32)
DEFINE_LOOPS(int64_t, i64)
DEFINE_LOOPS(uint8_t, u8)
DEFINE_LOOPS(uint16_t, u16)
DEFINE_LOOPS(uint32_t, u32)
DEFINE_LOOPS(uint64_t, u64)
DEFINE_LOOPS(float, f32)
DEFINE_LOOPS(double, f64)

/* =====================================================================================================
 * assign / assign_arbitary with cast (cpu_serial/assignment.rs:6-153; casts promotion.rs: Rust `as`)
 * ===================================================================================================== */
static inline long double load_as_ld(int t, const void *p, int64_t i, int *is_float, int64_t *iv, uint64_t *uv) {
    *is_float = 0;
    switch (t) {
        case T_BOOL: *uv = ((const uint8_t *)p)[i] != 0; *iv = (int64_t)*uv; return (long double)*uv;
        case T_I8: *iv = ((const int8_t *)p)[i]; *uv = (uint64_t)*iv; return (long double)*iv;
        case T_I16: *iv = ((const int16_t *)p)[i]; *uv = (uint64_t)*iv; return (long double)*iv;
        case T_I32: *iv = ((const int32_t *)p)[i]; *uv = (uint64_t)*iv; return (long double)*iv;
        case T_I64: *iv = ((const int64_t *)p)[i]; *uv = (uint64_t)*iv; return (long double)*iv;
        case T_U8: *uv = ((const uint8_t *)p)[i]; *iv = (int64_t)*uv; return (long double)*uv;
        case T_U16: *uv = ((const uint16_t *)p)[i]; *iv = (int64_t)*uv; return (long double)*uv;
        case T_U32: *uv = ((const uint32_t *)p)[i]; *iv = (int64_t)*uv; return (long double)*uv;
        case T_U64: *uv = ((const uint64_t *)p)[i]; *iv = (int64_t)*uv; return (long double)*uv;
        case T_F32: *is_float = 1; return (long double)((const float *)p)[i];
        case T_F64: *is_float = 1; return (long double)((const double *)p)[i];
    }
    return 0;
}

/* float -> integer `as`: saturating, NaN -> 0, truncation toward zero */
static inline int64_t sat_i(long double x, int64_t lo, int64_t hi) {
    if (x != x) return 0;
    if (x <= (long double)lo) return lo;
    if (x >= (long double)hi) return hi;
    return (int64_t)x;
}
static inline uint64_t sat_u(long double x, uint64_t hi) {
    if (x != x || x <= 0) return 0;
    if (x >= (long double)hi) return hi;
    return (uint64_t)x;
}

static void store_cast(int tc, void *c, int64_t ci, int ta, const void *a, int64_t ai) {
    int isf;
    int64_t iv = 0;
    uint64_t uv = 0;
    long double x = load_as_ld(ta, a, ai, &isf, &iv, &uv);
    int a_unsigned = (ta == T_U8 || ta == T_U16 || ta == T_U32 || ta == T_U64 || ta == T_BOOL);
    switch (tc) {
        case T_BOOL: ((uint8_t *)c)[ci] = isf ? (x != 0) : (a_unsigned ? uv != 0 : iv != 0); break;
        case T_I8: ((int8_t *)c)[ci] = isf ? (int8_t)sat_i(x, INT8_MIN, INT8_MAX) : (int8_t)(a_unsigned ? uv : (uint64_t)iv); break;
        case T_I16: ((int16_t *)c)[ci] = isf ? (int16_t)sat_i(x, INT16_MIN, INT16_MAX) : (int16_t)(a_unsigned ? uv : (uint64_t)iv); break;
        case T_I32: ((int32_t *)c)[ci] = isf ? (int32_t)sat_i(x, INT32_MIN, INT32_MAX) : (int32_t)(a_unsigned ? uv : (uint64_t)iv); break;
        case T_I64: ((int64_t *)c)[ci] = isf ? sat_i(x, INT64_MIN, INT64_MAX) : (int64_t)(a_unsigned ? uv : (uint64_t)iv); break;
        case T_U8: ((uint8_t *)c)[ci] = isf ? (uint8_t)sat_u(x, UINT8_MAX) : (uint8_t)(a_unsigned ? uv : (uint64_t)iv); break;
        case T_U16: ((uint16_t *)c)[ci] = isf ? (uint16_t)sat_u(x, UINT16_MAX) : (uint16_t)(a_unsigned ? uv : (uint64_t)iv); break;
        case T_U32: ((uint32_t *)c)[ci] = isf ? (uint32_t)sat_u(x, UINT32_MAX) : (uint32_t)(a_unsigned ? uv : (uint64_t)iv); break;
        case T_U64: ((uint64_t *)c)[ci] = isf ? sat_u(x, UINT64_MAX) : (a_unsigned ? uv : (uint64_t)iv); break;
        case T_F32:
            if (isf) ((float *)c)[ci] = (ta == T_F32) ? ((const float *)a)[ai] : (float)((const double *)a)[ai];
            else ((float *)c)[ci] = a_unsigned ? (float)uv : (float)iv;
            break;
        case T_F64:
            if (isf) ((double *)c)[ci] = (ta == T_F32) ? (double)((const float *)a)[ai] : ((const double *)a)[ai];
            else ((double *)c)[ci] = a_unsigned ? (double)uv : (double)iv;
            break;
    }
}

static size_t tsize(int t) {
    switch (t) {
        case T_BOOL: case T_I8: case T_U8: return 1;
        case T_I16: case T_U16: return 2;
        case T_I32: case T_U32: case T_F32: return 4;
        default: return 8;
    }
}

/* assign: same logical index (cpu_serial/assignment.rs:91-119) */
void orc_assign(int tc, void *c, const orc_layout *lc, int ta, const void *a, const orc_layout *la) {
    int64_t n = layout_size(lc);
    if (n == 0) return;
    orc_iter ic, ia;
    iter_init_at(&ic, lc, 0);
    iter_init_at(&ia, la, 0);
    for (int64_t k = 0; k < n; ++k) {
        if (tc == ta) memcpy((char *)c + ic.off * tsize(tc), (const char *)a + ia.off * tsize(ta), tsize(tc));
        else store_cast(tc, c, ic.off, ta, a, ia.off);
        iter_next(&ic); iter_next(&ia);
    }
}

/* assign_arbitary: k-th element of lc <- k-th element of la; lc/la arrive already translated to
 * col-major iteration (reverse_axes for a row-major device), cpu_serial/assignment.rs:39-67.
 * `parallel`: the rayon path (cpu_rayon/assignment.rs:41-93) -- single-thread linear copy when `contig`,
 * else the flattened range split over threads, each running two per-element odometers. */
void orc_assign_arbitary(int tc, void *c, const orc_layout *lc, int ta, const void *a, const orc_layout *la,
                         int contig, int parallel) {
    int64_t n = layout_size(lc);
    if (n == 0) return;
    if (contig) {
        if (tc == ta) {
            memmove((char *)c + lc->offset * tsize(tc), (const char *)a + la->offset * tsize(ta), (size_t)n * tsize(tc));
        } else {
            for (int64_t k = 0; k < n; ++k) store_cast(tc, c, lc->offset + k, ta, a, la->offset + k);
        }
        return;
    }
#pragma omp parallel if (parallel)
    {
        int nt = parallel ? orc_num_threads_in() : 1, tid = parallel ? orc_thread_id() : 0;
        int64_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
        orc_iter ic, ia;
        iter_init_at(&ic, lc, lo);
        iter_init_at(&ia, la, lo);
        if (tc == ta && tsize(tc) == 8) {
            uint64_t *cp = (uint64_t *)c;
            const uint64_t *ap = (const uint64_t *)a;
            for (int64_t k = lo; k < hi; ++k) { cp[ic.off] = ap[ia.off]; iter_next(&ic); iter_next(&ia); }
        } else if (tc == ta && tsize(tc) == 4) {
            uint32_t *cp = (uint32_t *)c;
            const uint32_t *ap = (const uint32_t *)a;
            for (int64_t k = lo; k < hi; ++k) { cp[ic.off] = ap[ia.off]; iter_next(&ic); iter_next(&ia); }
        } else {
            for (int64_t k = lo; k < hi; ++k) {
                if (tc == ta) memcpy((char *)c + ic.off * tsize(tc), (const char *)a + ia.off * tsize(ta), tsize(tc));
                else store_cast(tc, c, ic.off, ta, a, ia.off);
                iter_next(&ic); iter_next(&ia);
            }
        }
    }
}

/* broadcast fix-up of reduce_axes (cpu_serial/reduction.rs:330-349): replicate into stride-0 kept axes */
void orc_reduce_axes_bcast_fixup(int t, void *out, const orc_layout *l_o0, const orc_layout *l_ocd, int64_t offset) {
    int64_t n0 = layout_size(l_o0), n1 = layout_size(l_ocd);
    orc_iter i0;
    iter_init_at(&i0, l_o0, 0);
    for (int64_t p = 0; p < n0; ++p) {
        orc_iter i1;
        iter_init_at(&i1, l_ocd, 0);
        for (int64_t q = 0; q < n1; ++q) {
            int64_t dst = i0.off + i1.off - offset;
            memcpy((char *)out + dst * tsize(t), (const char *)out + i1.off * tsize(t), tsize(t));
            iter_next(&i1);
        }
        iter_next(&i0);
    }
}
