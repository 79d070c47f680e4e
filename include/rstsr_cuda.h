/*
 * rstsr_cuda.h -- C ABI of the B200-native `DeviceCuda` backend for rstsr.
 *
 * This is the drop-in boundary: one entry point per *trait method family* of the
 * reference's device-trait surface (citations are paths inside RESTGroup/rstsr
 * v0.7.10).  Plain pointers and sizes only; dtype and op are enums.  All
 * `void *` tensor arguments are DEVICE pointers to the start of the raw storage
 * (the reference's `Raw`), layouts address them in ELEMENTS (shape, stride, offset;
 * strides may be zero or negative), exactly like `Layout<D>`
 * (rstsr-common/src/layout/layoutbase.rs:15-23).
 *
 * Every function returns an `rc_status` (0 = ok).  The message of the last failure on
 * the calling thread is available from `rc_last_error()`; nothing aborts or throws
 * across the boundary.  The reference's `Result<T>` / `RSTSRError` kinds
 * (rstsr-common/src/error.rs:12-49) map onto `rc_status`.
 *
 * Ops are stream-ordered on the device handle's stream.  Entry points that return
 * host-visible data (`rc_reduce_all`, `rc_memcpy_d2h`, `rc_get_index`) synchronise
 * before returning, so -- like the reference -- every result is complete when the
 * caller can observe it.
 *
 * There is NO CPU fallback: if the CUDA runtime or a device is unavailable every
 * compute entry point fails with RC_ERR_DEVICE.
 *
 * Environment variables (read once per process; experiments only -- every default is the measured best and none
 * changes a result; DESIGN.md "Tuning knobs" has the sweeps behind them):
 *   RC_TILE_MIN_X / RC_TILE_MIN_Y, RC_TILE_RECT, RC_RECT_MAX_Y8 / _Y4, RC_RECT_X_LO8 / HI8 / LO4 / HI4   tile-kernel selection
 *   RC_TILE_BULK (0)      1: eligible 8-byte permuted copies take the TMA kernel (cp.async.bulk.tensor)
 *   RC_TILE_NARROW (1), RC_EW_OUTER (1), RC_SEL_WINDOW (1)   0: 1-/2-byte word tile, outer kernel, windowed gather off
 *   RC_ROWS_PER_CTA (1), RC_TCOL_MAX (64), RC_COLS_MIN (2), RC_TRI_TILE / RC_UNPACK_TILE (64)
 *   RC_COMM_PEER (1)      0: sharded reductions use ncclAllReduce for every size (no NVLink peer window)
 *   RC_COMM_TIMEOUT_S (120)   bound of the in-kernel wait for a lost rank in the peer-window exchange
 */
#ifndef RSTSR_CUDA_H
#define RSTSR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_MAX_NDIM 16

/* ---- status codes: mirror RSTSRError (rstsr-common/src/error.rs:12-49) ---- */
typedef enum rc_status {
    RC_OK = 0,
    RC_ERR_VALUE_OUT_OF_RANGE = 1,
    RC_ERR_INVALID_VALUE = 2,
    RC_ERR_INVALID_LAYOUT = 3,
    RC_ERR_RUNTIME = 4,
    RC_ERR_DEVICE_MISMATCH = 5,
    RC_ERR_UNIMPLEMENTED = 6,
    RC_ERR_MEMORY = 7,
    RC_ERR_DEVICE = 8, /* CUDA / NCCL failure -> RSTSRError::DeviceError(String) */
    RC_ERR_INDEX = 9
} rc_status;

/* ---- element types; codes are stable ---- */
typedef enum rc_dtype {
    RC_BOOL = 0, /* 1 byte, 0/1 (Rust `bool`) */
    RC_I8 = 1,
    RC_I16 = 2,
    RC_I32 = 3,
    RC_I64 = 4, /* also isize */
    RC_U8 = 5,
    RC_U16 = 6,
    RC_U32 = 7,
    RC_U64 = 8, /* also usize */
    RC_F32 = 9,
    RC_F64 = 10,
    /* round 2: half and complex element types (half::f16, half::bf16, num::Complex<f32>, num::Complex<f64>).
     * Arithmetic on the half types is f32 compute + ONE rounding, as the `half` crate does; complex mul / div use the
     * textbook formulas of num-complex.  Covered: storage, copy / to_contig / gather (raw words), fill, casts
     * half <-> f32 / f64 / bool / integers (through f64 / f32) and every primitive -> complex `(v as R, 0)`, c32 <-> c64;
     * operand promotion as in the reference's table (bool x T, complex x primitive, c32 x c64: promotion.rs:195-200,
     * :368-545; a half type pairs with itself and bool only, as there); + - * / neg, comparisons (== != only for complex),
     * maximum / minimum and the float math functions for half; abs / real / imag (real output), conj, square, reciprocal and
     * every ComplexFloat function of the reference for complex (exp log log2 log10 sqrt sin cos tan asin acos atan sinh
     * cosh tanh asinh acosh atanh; sign = z / |z|; is_nan / is_infinite / is_finite); elementwise isclose; reductions sum / prod / mean / var / std / l2_norm
     * (all four; var / std / l2_norm of complex are real), max / min / argmin / argmax / count_nonzero (half); vecdot
     * (sum conj(a) b), allclose_all; linspace (all four, in the type's own arithmetic), arange (half: the reference's
     * serial recurrence), tril / triu.
     * Anything else on these types is RC_ERR_UNIMPLEMENTED. */
    RC_F16 = 11,
    RC_BF16 = 12,
    RC_C32 = 13, /* 8 bytes: re, im as f32 */
    RC_C64 = 14  /* 16 bytes: re, im as f64 */
} rc_dtype;

/* FlagOrder (rstsr-common/src/flags.rs:60-87) */
typedef enum rc_order { RC_ROW_MAJOR = 0, RC_COL_MAJOR = 1 } rc_order;

/* TensorIterOrder subset accepted for copies (rstsr-common/src/flags.rs:92-136) */
typedef enum rc_iter_order { RC_ITER_C = 0, RC_ITER_F = 1, RC_ITER_A = 2, RC_ITER_K = 3 } rc_iter_order;

/* Layout<IxD>: rstsr-common/src/layout/layoutbase.rs:15-23 */
typedef struct rc_layout {
    int32_t ndim;
    int64_t shape[RC_MAX_NDIM];
    int64_t stride[RC_MAX_NDIM]; /* in elements; 0 = broadcast, < 0 = flipped */
    int64_t offset;              /* in elements, >= 0 */
} rc_layout;

/* "ternary" ops c = a o b: Op{Add..Shr}API (rstsr-core/src/operators/ops/op_ternary_arithmetic.rs:16-48)
 * and the binary-function traits (rstsr-core/src/operators/ops/op_ternary_common.rs:3-102). */
typedef enum rc_binop {
    RC_ADD = 0,
    RC_SUB = 1,
    RC_MUL = 2,
    RC_DIV = 3,
    RC_REM = 4,
    RC_BITOR = 5,
    RC_BITAND = 6,
    RC_BITXOR = 7,
    RC_SHL = 8,
    RC_SHR = 9,
    /* same-type function ops */
    RC_MAXIMUM = 10,
    RC_MINIMUM = 11,
    RC_FLOOR_DIVIDE = 12,
    RC_POW = 13,
    RC_ATAN2 = 14,
    RC_COPYSIGN = 15,
    RC_HYPOT = 16,
    RC_LOGADDEXP = 17,
    RC_NEXTAFTER = 18,
    /* comparison ops: output dtype is RC_BOOL */
    RC_EQ = 32,
    RC_NE = 33,
    RC_LT = 34,
    RC_LE = 35,
    RC_GT = 36,
    RC_GE = 37
} rc_binop;

/* unary ops a = f(b): OpNeg/NotAPI (ops/op_binary_arithmetic.rs:90-109) and the unary
 * math traits (ops/op_binary_common.rs:3-57). */
typedef enum rc_unop {
    RC_NEG = 0,
    RC_NOT = 1,
    RC_ABS = 2,
    RC_SQUARE = 3,
    RC_SIGN = 4,
    RC_SQRT = 5,
    RC_EXP = 6,
    RC_EXPM1 = 7,
    RC_LOG = 8,
    RC_LOG2 = 9,
    RC_LOG10 = 10,
    RC_SIN = 11,
    RC_COS = 12,
    RC_TAN = 13,
    RC_ASIN = 14,
    RC_ACOS = 15,
    RC_ATAN = 16,
    RC_SINH = 17,
    RC_COSH = 18,
    RC_TANH = 19,
    RC_ASINH = 20,
    RC_ACOSH = 21,
    RC_ATANH = 22,
    RC_FLOOR = 23,
    RC_CEIL = 24,
    RC_ROUND = 25,
    RC_TRUNC = 26,
    RC_RECIPROCAL = 27, /* also OpInvAPI */
    RC_CONJ = 28,       /* identity on real types */
    RC_REAL = 29,       /* identity on real types */
    RC_IMAG = 30,       /* zero on real types */
    /* boolean-output predicates: output dtype is RC_BOOL */
    RC_ISNAN = 48,
    RC_ISINF = 49,
    RC_ISFINITE = 50,
    RC_SIGNBIT = 51
} rc_unop;

/* reductions: Op{Sum,Min,Max,Prod,Mean}API (rstsr-core/src/operators/reduction.rs:3-33) */
typedef enum rc_redop {
    RC_SUM = 0, RC_PROD = 1, RC_MAX = 2, RC_MIN = 3, RC_MEAN = 4,
    /* "next" reductions (same kernels, other monoids; operators/reduction.rs:35-95): */
    RC_VAR = 5,            /* population variance q/n - (s/n)^2 (auto_impl/reduction.rs:207-262) */
    RC_STD = 6,
    RC_L2_NORM = 7,
    RC_ARGMIN = 8,         /* output u64: row-major index within the reduced axes (in the order given) */
    RC_ARGMAX = 9,
    RC_ALL = 10,           /* bool -> bool */
    RC_ANY = 11,
    RC_COUNT_NONZERO = 12  /* any dtype -> u64; on bool this is OpSumBoolAPI */
} rc_redop;

typedef struct rc_device rc_device; /* opaque: {ordinal, default_order, stream, workspaces} */

/* ------------------------------------------------------------------------------------------
 * Library / error plumbing
 * ---------------------------------------------------------------------------------------- */
const char *rc_last_error(void);  /* thread-local message of the last non-OK status */
const char *rc_version(void);     /* "rstsr-cuda <semver> sm_100a" */
int rc_device_count(int *count);  /* number of visible CUDA devices */

/* ------------------------------------------------------------------------------------------
 * DeviceBaseAPI (rstsr-core/src/storage/device.rs:3-7; DeviceFaer impl device_faer/device.rs:46-60)
 * ---------------------------------------------------------------------------------------- */
int rc_device_create(int ordinal, rc_order default_order, rc_device **out);
/* same, but ops are enqueued on a caller-owned cudaStream_t (e.g. torch's current stream) */
int rc_device_create_on_stream(int ordinal, rc_order default_order, void *cuda_stream, rc_device **out);
int rc_device_destroy(rc_device *dev);
int rc_device_default_order(const rc_device *dev, rc_order *out);
int rc_device_set_default_order(rc_device *dev, rc_order order);
/* same GPU, same default order AND same stream: two handles with different streams are not ordered against each
 * other, so storage may only move between them through rc_memcpy_peer / rc_device_wait (cf. DeviceFaer::same_device =
 * equal pool size & order, device_faer/device.rs:46-51) */
int rc_device_same_device(const rc_device *a, const rc_device *b, int *same);
/* NUMA node of the GPU's PCIe root (from /sys/bus/pci/devices/<bus id>/numa_node), -1 if unknown */
int rc_device_numa_node(const rc_device *dev, int *node);
int rc_device_ordinal(const rc_device *dev, int *ordinal);
int rc_device_stream(const rc_device *dev, void **cuda_stream);
int rc_device_synchronize(rc_device *dev);
/* number of kernels this handle has launched (bench.py's `gpu_launches`) */
int rc_device_launch_count(const rc_device *dev, uint64_t *count);

/* ------------------------------------------------------------------------------------------
 * DeviceRawAPI / DeviceStorageAPI / DeviceCreationAnyAPI
 * (storage/device.rs:9-54, storage/creation.rs:3-39, auto_impl/creation.rs:5-72)
 *   uninit_impl/empty_impl -> rc_malloc          outof_cpu_vec/from_cpu_vec -> rc_malloc + rc_memcpy_h2d
 *   to_cpu_vec/into_cpu_vec -> rc_memcpy_d2h     get_index/set_index -> rc_get_index/rc_set_index
 *   zeros_impl -> rc_malloc + rc_memset          full_impl/ones_impl -> rc_malloc + rc_fill
 *   Raw::clone -> rc_malloc + rc_memcpy_d2d      Raw::drop -> rc_free
 * ---------------------------------------------------------------------------------------- */
int rc_malloc(rc_device *dev, size_t nbytes, void **out);
int rc_free(rc_device *dev, void *ptr);
int rc_memcpy_h2d(rc_device *dev, void *dst_dev, const void *src_host, size_t nbytes);
int rc_memcpy_d2h(rc_device *dev, void *dst_host, const void *src_dev, size_t nbytes); /* synchronises */
int rc_memcpy_d2d(rc_device *dev, void *dst_dev, const void *src_dev, size_t nbytes);
int rc_memset(rc_device *dev, void *dst_dev, int byte, size_t nbytes);
/* Stream-ordered transfers for pipelined staging (outof_cpu_vec / to_cpu_vec of large tensors): nothing
 * synchronises; the host buffer must be pinned (rc_host_alloc) and stay alive until rc_device_synchronize.
 * The 2-D form moves `height` rows of `width_bytes` with independent pitches (a strided slab of a tensor). */
int rc_memcpy_h2d_async(rc_device *dev, void *dst_dev, const void *src_host, size_t nbytes);
int rc_memcpy_d2h_async(rc_device *dev, void *dst_host, const void *src_dev, size_t nbytes);
int rc_memcpy2d_d2h_async(rc_device *dev, void *dst_host, size_t dst_pitch, const void *src_dev, size_t src_pitch,
                          size_t width_bytes, size_t height);
int rc_memcpy2d_h2d_async(rc_device *dev, void *dst_dev, size_t dst_pitch, const void *src_host, size_t src_pitch,
                          size_t width_bytes, size_t height);
/* Cross-handle ordering: everything enqueued on `waiter` after this call runs after everything enqueued on
 * `signaler` before it (cudaEventRecord + cudaStreamWaitEvent).  Handles of one ordinal are "the same device"
 * (rc_device_same_device); using several of them is how copies overlap kernels. */
int rc_device_wait(rc_device *waiter, rc_device *signaler);
/* DeviceChangeAPI between two DeviceCuda handles (rstsr-core/src/storage/conversion.rs:3-21; pattern
 * crates-device/rstsr-openblas/src/conversion.rs:3-48): copy `nbytes` from `src` on src_dev to `dst` on dst_dev
 * (same or different GPU; NVLink peer copy when available).  Stream-ordered on both handles: runs after the work
 * already enqueued on src_dev, before anything enqueued later on either handle.  CPU <-> CUDA is
 * rc_memcpy_h2d / rc_memcpy_d2h. */
int rc_memcpy_peer(rc_device *dst_dev, void *dst_ptr, rc_device *src_dev, const void *src_ptr, size_t nbytes);
int rc_get_index(rc_device *dev, rc_dtype dtype, const void *a, int64_t index, void *host_out);
int rc_set_index(rc_device *dev, rc_dtype dtype, void *a, int64_t index, const void *host_value);
/* pinned host staging buffers for outof_cpu_vec / to_cpu_vec */
int rc_host_alloc(size_t nbytes, void **out);
/* same, with the pages placed on NUMA node `node` before they are pinned: mbind(MPOL_BIND) (*bound_out = 1), or --
 * where the container filters mbind -- first touch from a CPU of that node (*bound_out = 2); node < 0 = no placement;
 * a failed placement is not an error: the buffer is still pinned, *bound_out = 0.  Staging buffers on the GPU's own
 * socket keep H2D / D2H copies off the inter-socket link.  Free with rc_host_free. */
int rc_host_alloc_on_node(size_t nbytes, int node, void **out, int *bound_out);
int rc_host_free(void *ptr);
size_t rc_dtype_size(rc_dtype dtype);

/* ------------------------------------------------------------------------------------------
 * Host-side layout algebra: the parts of L0/L4 that decide what the device is asked to do and
 * which layout the result has.  Reference-identical by contract.
 * ---------------------------------------------------------------------------------------- */
/* Layout::new validity: bounds_index + check_strides(skip_zero=true) (layoutbase.rs:237-320,396-404) */
int rc_layout_check(const rc_layout *l);
int rc_layout_bounds_index(const rc_layout *l, int64_t *min_out, int64_t *max_out);
int rc_layout_c_contig(const rc_layout *l, int *out);
int rc_layout_f_contig(const rc_layout *l, int *out);
/* shape.c() / shape.f()  (layoutbase.rs:574-601) */
int rc_layout_new_contig(const int64_t *shape, int ndim, rc_order order, int64_t offset, rc_layout *out);
/* broadcast_layout (rstsr-common/src/layout/broadcast.rs:166-182) */
int rc_layout_broadcast(const rc_layout *la, const rc_layout *lb, rc_order order, rc_layout *la_out,
                        rc_layout *lb_out);
/* get_layout_for_binary_op (rearrangement.rs:394-459): output layout of `a o b` (K iteration order) */
int rc_layout_for_binary_op(const rc_layout *la, const rc_layout *lb, rc_order order, rc_layout *lc_out);
/* layout_for_array_copy (rearrangement.rs:125-152): output layout of unary ops, scalar ops, to_owned */
int rc_layout_for_array_copy(const rc_layout *la, rc_iter_order iter_order, rc_order default_order,
                             rc_layout *lc_out);
/* output layout of `*_axes` reductions: dim_split_axes (indexer.rs:453-478) + layout_for_array_copy(K)
 * (cpu_rayon/reduction.rs:147-153). `axes` may be negative; duplicates are an error (axis_index.rs:379-414). */
int rc_layout_for_reduce(const rc_layout *la, const int64_t *axes, int naxes, rc_layout *lo_out);
/* layout_reshapeable (rstsr-common/src/layout/reshape.rs:216-226): *viewable = 1 and *out set when the
 * reshape is a pure view; *viewable = 0 when a copy (rc_assign_arbitary) is needed. */
int rc_layout_reshapeable(const rc_layout *la, const int64_t *shape, int ndim, rc_order order, int *viewable,
                          rc_layout *out);
/* Layout::eq (layoutbase.rs:547-572) -- used by to_layout to decide view-vs-copy (to_layout.rs:20) */
int rc_layout_equal(const rc_layout *a, const rc_layout *b, int *equal);

/* ------------------------------------------------------------------------------------------
 * OpAssignAPI / OpAssignArbitaryAPI (rstsr-core/src/operators/assignment.rs:5-53;
 * auto_impl/assignment.rs:3-49; loops cpu_rayon/assignment.rs:14-225)
 * ---------------------------------------------------------------------------------------- */
/* c[idx] = cast(a[idx]); lc and la have the same shape (already broadcast) */
int rc_assign(rc_device *dev, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta, const void *a,
              const rc_layout *la);
/* k-th element of lc <- k-th element of la, k counted in the device's default order (C for row-major,
 * F for col-major); shapes may differ, sizes must match.  (Reference spelling "arbitary" kept.) */
int rc_assign_arbitary(rc_device *dev, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta, const void *a,
                       const rc_layout *la);
/* same with the pairing order given explicitly instead of taken from the handle (reshape / to_layout with an order
 * argument must not mutate a handle other threads share) */
int rc_assign_arbitary_order(rc_device *dev, rc_order order, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta,
                             const void *a, const rc_layout *la);
/* c[idx] = cast(fill); *fill is a host scalar of dtype tf */
int rc_fill(rc_device *dev, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype tf, const void *fill);

/* ------------------------------------------------------------------------------------------
 * DeviceCreationArangeAPI / ComplexFloatAPI (linspace) / TriAPI (rstsr-core/src/storage/creation.rs:41-62;
 * auto_impl/creation.rs:74-119; loops cpu_rayon/creation.rs:8-131, cpu_serial/op_tri.rs:524-590).
 * arange / linspace: the CALLEE allocates (free with rc_free); scalars are host values of `dtype`.
 * arange length and values follow the reference: floats via f64 arithmetic start + i*step, n = ceil((end-start)/step)
 * with a last element on/past `end` dropped; integers via isize arithmetic.  tril / triu zero the other triangle
 * of the last two axes IN PLACE (k-th diagonal kept), batched over the leading axes.
 * ---------------------------------------------------------------------------------------- */
int rc_arange(rc_device *dev, rc_dtype dtype, const void *start, const void *end, const void *step, void **out_dev,
              int64_t *n_out);
int rc_linspace(rc_device *dev, rc_dtype dtype, const void *start, const void *end, int64_t n, int endpoint,
                void **out_dev);
int rc_tril(rc_device *dev, rc_dtype dtype, void *a, const rc_layout *la, int64_t k);
int rc_triu(rc_device *dev, rc_dtype dtype, void *a, const rc_layout *la, int64_t k);

/* ------------------------------------------------------------------------------------------
 * Elementwise (auto_impl/op_ternary_arithmetic.rs, op_ternary_common.rs, op_binary_arithmetic.rs,
 * op_binary_common.rs; loops cpu_rayon/op_with_func.rs:13-390).
 * `dtype` is the operand type TA = TB; the output type is `dtype`, except RC_BOOL for comparisons /
 * predicates.  All layouts have the same shape (already broadcast by the caller).
 * ---------------------------------------------------------------------------------------- */
int rc_op_mutc_refa_refb(rc_device *dev, rc_binop op, rc_dtype dtype, void *c, const rc_layout *lc, const void *a,
                         const rc_layout *la, const void *b, const rc_layout *lb);
int rc_op_mutc_refa_numb(rc_device *dev, rc_binop op, rc_dtype dtype, void *c, const rc_layout *lc, const void *a,
                         const rc_layout *la, const void *b_host_scalar);
int rc_op_mutc_numa_refb(rc_device *dev, rc_binop op, rc_dtype dtype, void *c, const rc_layout *lc,
                         const void *a_host_scalar, const void *b, const rc_layout *lb);
/* in-place: a = a o b  (Op*AssignAPI, OpLConsume*API);  reverse != 0: a = b o a  (OpRConsume*API) */
int rc_op_muta_refb(rc_device *dev, rc_binop op, rc_dtype dtype, void *a, const rc_layout *la, const void *b,
                    const rc_layout *lb, int reverse);
int rc_op_muta_numb(rc_device *dev, rc_binop op, rc_dtype dtype, void *a, const rc_layout *la,
                    const void *b_host_scalar, int reverse);
/* unary with output a = f(b) (dtype = type of b; output dtype per op) and in-place a = f(a) */
int rc_unary_muta_refb(rc_device *dev, rc_unop op, rc_dtype dtype, void *a, const rc_layout *la, const void *b,
                       const rc_layout *lb);
int rc_unary_muta(rc_device *dev, rc_unop op, rc_dtype dtype, void *a, const rc_layout *la);
/* output dtype of a reduction (u64 for arg* and count_nonzero, bool for all/any, else the element type) */
int rc_redop_out_dtype(rc_redop op, rc_dtype dtype, rc_dtype *out);
/* output dtype of an elementwise op for operand dtype `dtype` (bool for comparisons/predicates) */
int rc_binop_out_dtype(rc_binop op, rc_dtype dtype, rc_dtype *out);

/* ------------------------------------------------------------------------------------------
 * Mixed operand types: the 15 binary-function traits are generic over <TA, TB> and promote
 * (rstsr-core/src/operators/ops/op_ternary_common.rs:22-57; impls
 * rstsr-core/src/feature_rayon/auto_impl/op_ternary_common.rs:6-120):
 *   R = DTypePromoteAPI<TB>::Res of TA   (NumPy's table, rstsr-dtype-traits/src/promotion.rs:186-300:
 *                                          i32 x f32 -> f64, i8 x u8 -> i16, i64 x u64 -> f64, bool x T -> T ...;
 *                                          :368-545: c32 x i8/i16/u8/u16/f32 -> c32, c32 x wider -> c64, c64 x any -> c64)
 *   atan2 copysign hypot nextafter logaddexp : both operands -> R -> DTypeIntoFloatAPI (ints -> f64), TOut = that float
 *   maximum minimum floor_divide             : both -> R, TOut = R
 *   == != < <= > >=                          : both -> R, TOut = bool
 *   pow                                      : num::Pow<TB> for TA: float x same float (powf), float x i8/u8/i16/u16/i32
 *                                              (powi), int x u8/u16/u32/u64 (wrapping power); TOut = TA
 *   + - * / % | & ^ << >>                    : both -> R, TOut = R (the reference requires TA: Op<TB>; its tensor layer
 *                                              promotes first)
 * `tc` must be rc_binop_out_dtype_ex(op, ta, tb), else RC_ERR_INVALID_VALUE.  ta == tb runs the single fused kernel
 * of rc_op_mutc_refa_refb; + - * / on f32 / i32 / i64 with f64 and i32 with i64 widen in registers inside one kernel;
 * otherwise the operand(s) whose type differs from the compute type are cast first (element-exact `as` casts,
 * broadcast axes kept compact), then the same kernel runs.  All three are bit-identical to promote_pair +
 * into_float + f per element.
 * ---------------------------------------------------------------------------------------- */
int rc_dtype_promote(rc_dtype ta, rc_dtype tb, rc_dtype *out);
int rc_binop_out_dtype_ex(rc_binop op, rc_dtype ta, rc_dtype tb, rc_dtype *out);
int rc_op_mutc_refa_refb_ex(rc_device *dev, rc_binop op, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta,
                            const void *a, const rc_layout *la, rc_dtype tb, const void *b, const rc_layout *lb);
int rc_op_mutc_refa_numb_ex(rc_device *dev, rc_binop op, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta,
                            const void *a, const rc_layout *la, rc_dtype tb, const void *b_host_scalar);
int rc_op_mutc_numa_refb_ex(rc_device *dev, rc_binop op, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta,
                            const void *a_host_scalar, rc_dtype tb, const void *b, const rc_layout *lb);
/* OpIsCloseAPI (rstsr-core/src/operators/ops/op_ternary_common.rs:59-102; isclose of
 * rstsr-dtype-traits/src/isclose.rs:92-106 with TE = f64): c (bool) = |a-b| <= atol + rtol*|b| || (equal_nan && both NaN) */
int rc_isclose(rc_device *dev, rc_dtype dtype, void *c_bool, const rc_layout *lc, const void *a, const rc_layout *la,
               const void *b, const rc_layout *lb, double rtol, double atol, int equal_nan);
int rc_unop_out_dtype(rc_unop op, rc_dtype dtype, rc_dtype *out);

/* ------------------------------------------------------------------------------------------
 * Reductions: Op{Sum,Prod,Max,Min,Mean}API (rstsr-core/src/operators/reduction.rs:3-33;
 * auto_impl/reduction.rs:7-205; loops cpu_rayon/reduction.rs:20-328)
 * ---------------------------------------------------------------------------------------- */
/* `*_all` -> host scalar of type `dtype`.  max/min of a zero-size layout is RC_ERR_INVALID_VALUE. */
int rc_reduce_all(rc_device *dev, rc_redop op, rc_dtype dtype, const void *a, const rc_layout *la, void *host_out);
/* same, but the scalar stays on the device (dev_out: 1 element) and nothing synchronises: the
 * per-GPU partial of a sharded reduction, to be combined by rc_comm_all_reduce. */
int rc_reduce_all_device(rc_device *dev, rc_redop op, rc_dtype dtype, const void *a, const rc_layout *la,
                         void *dev_out);
/* `*_axes`: the CALLEE allocates the output (free with rc_free) and chooses its layout
 * (operators/reduction.rs:26-32): *out_dev has lo_out->size() elements, lo_out == rc_layout_for_reduce(). */
int rc_reduce_axes(rc_device *dev, rc_redop op, rc_dtype dtype, const void *a, const rc_layout *la,
                   const int64_t *axes, int naxes, void **out_dev, rc_layout *lo_out);
/* same into caller-provided storage with an explicit output layout (shape = kept axes of la) */
int rc_reduce_axes_into(rc_device *dev, rc_redop op, rc_dtype dtype, const void *a, const rc_layout *la,
                        const int64_t *axes, int naxes, void *out_dev, const rc_layout *lo);

/* OpUnraveledArgMin/MaxAPI (rstsr-core/src/operators/reduction.rs:35-55; reduce_*_unraveled_arg_cpu_serial,
 * rstsr-native-impl/src/cpu_serial/reduction.rs:421-529): the position of the first extreme element as an index TUPLE.
 *   _all:  index_out[0 .. la->ndim) = the position within `la` (host memory; synchronises).
 *   _axes: the callee allocates u64[lo_out->size()][naxes] (free with rc_free): for the output element that lives at
 *          element offset m of lo_out (layout as rc_layout_for_reduce), entries m * naxes + k, k < naxes, are its
 *          position within the REDUCED axes in the order given (the reference returns Vec<IxD> elements; a device
 *          buffer holds them as fixed-length u64 tuples).
 * op is RC_ARGMIN or RC_ARGMAX; a zero-size input is RC_ERR_INVALID_LAYOUT as in the reference. */
int rc_reduce_unraveled_arg_all(rc_device *dev, rc_redop op, rc_dtype dtype, const void *a, const rc_layout *la,
                                int64_t *index_out);
int rc_reduce_unraveled_arg_axes(rc_device *dev, rc_redop op, rc_dtype dtype, const void *a, const rc_layout *la,
                                 const int64_t *axes, int naxes, void **out_dev, rc_layout *lo_out);

/* ------------------------------------------------------------------------------------------
 * Binary reductions: two input streams folded in one pass.
 *   rc_vecdot:       DeviceVecdotAPI::vecdot (rstsr-core/src/device_cpu_serial/linalg/vecdot.rs:4-29,
 *                    rstsr-native-impl/src/cpu_serial/vecdot.rs:6-168): c[m] = sum_s a[m,s] * b[m,s];
 *                    `axes_a` / `axes_b` pair the contracted axes (dim_split_axes order), the remaining
 *                    axes of a and b must broadcast to lc's shape in the handle's default order.
 *                    Errors as the reference: contracted shapes differ / c not broadcast from a, b
 *                    -> RC_ERR_INVALID_LAYOUT.
 *   rc_allclose_all: OpAllCloseAPI::allclose_all (rstsr-core/src/device_cpu_serial/reduction.rs:660-683)
 *                    with isclose of rstsr-dtype-traits/src/isclose.rs:92-106 and TE = f64:
 *                    all(|a-b| <= atol + rtol*|b| || (equal_nan && a,b NaN)); la, lb already broadcast to
 *                    one shape (tensor level: rstsr-core/src/tensor/reduction.rs:324-351).  Zero size
 *                    -> RC_ERR_INVALID_VALUE.  Synchronises the stream; *result = 0 / 1.
 * Element types: f32, f64, i32, u32, i64, u64 (a, b and c share one dtype).
 * ---------------------------------------------------------------------------------------- */
int rc_vecdot(rc_device *dev, rc_dtype dtype, void *c, const rc_layout *lc, const void *a, const rc_layout *la,
              const void *b, const rc_layout *lb, const int64_t *axes_a, const int64_t *axes_b, int naxes);
int rc_allclose_all(rc_device *dev, rc_dtype dtype, const void *a, const rc_layout *la, const void *b,
                    const rc_layout *lb, double rtol, double atol, int equal_nan, int *result);

/* ------------------------------------------------------------------------------------------
 * Index-driven data movement.
 *   rc_index_select: DeviceIndexSelectAPI::index_select (rstsr-core/src/device_cpu_serial/adv_indexing.rs:3-20,
 *                    rstsr-native-impl/src/cpu_serial/adv_indexing.rs:3-90): c[.., i, ..] = a[.., indices[i], ..]
 *                    along `axis`; `indices` is a HOST array (the trait takes &[usize]); lc.shape[axis] must
 *                    equal n_indices (RC_ERR_INVALID_LAYOUT "Invalid index length."), an index outside
 *                    0..la.shape[axis] is RC_ERR_INDEX "Index out of range.".
 *   rc_pack_tri:     OpPackTriAPI::pack_tri (rstsr-core/src/device_cpu_serial/operators/op_tri.rs:4-28): `a` is the
 *                    PACKED output (.., n(n+1)/2), `b` the full input (.., n, n); the triangle is walked row by
 *                    row (rstsr-native-impl/src/cpu_serial/op_tri.rs:7-71).  Col-major handles see the axes
 *                    reversed ((n, n, ..) / (n(n+1)/2, ..)) and the triangle flipped, as the reference does.
 *   rc_unpack_tri:   OpUnpackTriAPI::unpack_tri (op_tri.rs:30-53): `a` is the FULL output, `b` the packed input;
 *                    symm: Sy/He mirror, Ay/Ah mirror with a minus sign and a ZERO diagonal, N leaves the other
 *                    triangle untouched (cpu_serial/op_tri.rs:152-446).  f32 / f64 (ComplexFloat in the reference).
 * ---------------------------------------------------------------------------------------- */
typedef enum rc_uplo { RC_UPLO_U = 0, RC_UPLO_L = 1 } rc_uplo;                                 /* FlagUpLo */
typedef enum rc_symm { RC_SYMM_SY = 0, RC_SYMM_HE = 1, RC_SYMM_AY = 2, RC_SYMM_AH = 3, RC_SYMM_N = 4 } rc_symm; /* FlagSymm */
int rc_index_select(rc_device *dev, rc_dtype dtype, void *c, const rc_layout *lc, const void *a, const rc_layout *la,
                    int axis, const int64_t *indices, int64_t n_indices);
int rc_pack_tri(rc_device *dev, rc_dtype dtype, void *a, const rc_layout *la, const void *b, const rc_layout *lb,
                rc_uplo uplo);
int rc_unpack_tri(rc_device *dev, rc_dtype dtype, void *a, const rc_layout *la, const void *b, const rc_layout *lb,
                  rc_uplo uplo, rc_symm symm);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: one process (or handle) per GPU; shards are independent except for reductions whose
 * sharded axis is reduced.  The collective is NCCL all-reduce over NVLink.
 * ---------------------------------------------------------------------------------------- */
typedef struct rc_comm rc_comm;
#define RC_COMM_ID_BYTES 128
int rc_comm_get_unique_id(uint8_t id[RC_COMM_ID_BYTES]);
int rc_comm_init_rank(rc_device *dev, int nranks, int rank, const uint8_t id[RC_COMM_ID_BYTES], rc_comm **out);
int rc_comm_destroy(rc_comm *comm);
/* *nranks, *rank of the communicator; *peer_window = 1 when results up to 256 KiB are combined through the
 * NVLink peer window (one kernel: fold of the local partial states + stores into every rank's window + rank-ordered
 * fold; bitwise identical on all ranks and run to run) instead of ncclAllReduce.  RC_COMM_PEER=0 in the environment
 * disables the window (NCCL only); RC_COMM_TIMEOUT_S bounds the in-kernel wait for a lost rank (default 120 s,
 * then the kernel traps instead of hanging the job).  Any out pointer may be NULL. */
int rc_comm_info(const rc_comm *comm, int *nranks, int *rank, int *peer_window);
/* Switch the peer window off (NCCL for every size) or back on.  COLLECTIVE in spirit: every rank must make the same
 * call between the same two reductions, otherwise the next reduction pairs a window kernel with an NCCL kernel and
 * times out.  Enabling fails with RC_ERR_DEVICE when the window could not be mapped at rc_comm_init_rank. */
int rc_comm_set_peer_window(rc_comm *comm, int enable);
/* in-place all-reduce of `count` elements with the reduction's combiner (sum/prod/max/min; mean = sum,
 * the caller divides by the global count via rc_op_muta_numb) on the device's stream.  All element types
 * (16-bit integers have no NCCL type: they go through the window, or widened to 32 bits above its size). */
int rc_comm_all_reduce(rc_comm *comm, rc_redop op, rc_dtype dtype, void *buf_dev, size_t count);
/* sharded `*_all` (reduce_all_cpu_rayon, rstsr-native-impl/src/cpu_rayon/reduction.rs:20-106, with the closures of
 * rstsr-core/src/feature_rayon/auto_impl/reduction.rs:14-63): local partial + cross-GPU combine + host scalar;
 * n_global is the GLOBAL element count (mean = combined sum / n_global; max / min of n_global == 0 is
 * RC_ERR_INVALID_VALUE on every rank).  A rank whose shard is empty contributes the monoid identity. */
int rc_reduce_all_sharded(rc_device *dev, rc_comm *comm, rc_redop op, rc_dtype dtype, const void *a,
                          const rc_layout *la, int64_t n_global, void *host_out);
/* sharded `*_axes` where the SHARDED axis is among the reduced ones (reduce_axes_cpu_rayon,
 * rstsr-native-impl/src/cpu_rayon/reduction.rs:109-328): `la` is this rank's shard, (out_dev, lo) the full-size
 * output (identical layout on every rank, shape = kept axes); every rank ends with the complete result.
 * n_reduced_global = product of the reduced extents of the UNSHARDED tensor (the mean's divisor).
 * sum / prod / max / min / mean. */
int rc_reduce_axes_sharded(rc_device *dev, rc_comm *comm, rc_redop op, rc_dtype dtype, const void *a,
                           const rc_layout *la, const int64_t *axes, int naxes, int64_t n_reduced_global, void *out_dev,
                           const rc_layout *lo);

#ifdef __cplusplus
}
#endif
#endif /* RSTSR_CUDA_H */
