// rc_ops.hpp -- typed dispatch entry points implemented by the .cu translation units.
#pragma once
#include "rc_canon.hpp"
#include "rc_device.hpp"

namespace rc {

struct EwArgs;

// c = a o b (arithmetic: rc_ew_arith.cu, bit ops: rc_ew_bit.cu, functions: rc_ew_func.cu, comparisons: rc_ew_cmp.cu)
void run_binary_arith(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
// promote_pair fused into the kernel for the common mixed pairs (rc_ew_mixed.cu); false = no such kernel
bool run_binary_promoted(rc_device *dev, rc_binop op, rc_dtype k, rc_dtype ta, rc_dtype tb, const CanonEw &c, const EwArgs &args);
void run_binary_bit(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
void run_binary_func(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
void run_binary_cmp(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
// pow with a mixed exponent (i32 for a float base, u32 for an integer base) and elementwise isclose (rc_ew_func.cu)
void run_binary_pow_mixed(rc_device *dev, rc_dtype t, const CanonEw &c, const EwArgs &args);
void run_isclose(rc_device *dev, rc_dtype t, const CanonEw &c, const EwArgs &args);
// Rust `as` between host scalars (rc_api.cu): 8 bytes out
void cast_host_scalar(rc_dtype tc, rc_dtype tf, const void *src, void *out8);
// extended element types f16 / bf16 / c32 / c64 (rc_ew_ext.cu, rc_reduce_extx.cu); the bool functions return false
// when the op / cast does not exist for the type
bool run_binary_ext(rc_device *dev, rc_binop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
bool run_unary_ext(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
bool run_cast_ext(rc_device *dev, rc_dtype tc, rc_dtype ta, const CanonEw &c, const EwArgs &args);
void run_copy16(rc_device *dev, const CanonEw &c, const EwArgs &args);
void run_fill16(rc_device *dev, const CanonEw &c, const EwArgs &args);
void host_to_ext(rc_dtype tc, double re, double im, void *out16);
void host_from_ext(rc_dtype t, const void *src, double *re, double *im);
void run_reduce_extx(rc_device *dev, rc_redop op, rc_dtype t, const CanonRed &cr, const void *a, void *out, int64_t n);
// a = f(b)
void run_unary(rc_device *dev, rc_unop op, rc_dtype t, const CanonEw &c, const EwArgs &args);
// c = cast(a) (rc_copy.cu); same dtype moves raw bits
void run_cast(rc_device *dev, rc_dtype tc, rc_dtype ta, const CanonEw &c, const EwArgs &args);
// c = value (already of dtype tc, host memory)
void run_fill(rc_device *dev, rc_dtype tc, const CanonEw &c, void *c_ptr, const void *value_tc);
// generic flattened-order copy when two shapes have no common refinement (rc_copy.cu)
void run_assign_arbitrary_generic(rc_device *dev, rc_dtype tc, void *c, const Layout &lc, rc_dtype ta, const void *a,
                                  const Layout &la, rc_order order);

// reductions (rc_reduce.cu): out has n_out elements laid out by `cr`
void run_reduce(rc_device *dev, rc_redop op, rc_dtype t, const CanonRed &cr, const void *a, void *out,
                int64_t mean_count);

// reduction of `a` over `axes` into (out, lo) WITHOUT taking dev->ws_mu (rc_api.cu): the caller holds the lock and
// owns dev->preq (see rc_device.hpp) -- used by the sharded reductions of rc_comm.cu
void reduce_local(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const Layout &la, const std::vector<int> &axes,
                  void *out, const Layout &lo);
rc_dtype redop_out_dtype(rc_redop op, rc_dtype t);

}  // namespace rc
