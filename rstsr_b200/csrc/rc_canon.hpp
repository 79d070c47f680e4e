// rc_canon.hpp -- host canonicaliser: turns reference layouts into the compact descriptors the kernels
// take by value.  This supersedes translate_to_col_major / translate_to_col_major_with_contig
// (rstsr-common/src/layout/rearrangement.rs:232-322): it also drops extent-1 axes, flips axes the
// output walks backwards, and merges every pair of axes that is jointly contiguous in all operands.
#pragma once
#include "rc_common.hpp"

namespace rc {

constexpr int KMAXD = 8;   // dims a kernel decomposes an index into (after merging)
constexpr int KMAXOPS = 3; // operands of one elementwise launch (output first)

// Same-shape operands, canonical axis order: dim 0 is the output's fastest axis.
struct CanonEw {
    int nops = 0;
    int ndim = 0;                        // >= 1 unless empty
    bool empty = false;                  // zero elements: nothing to launch
    std::vector<int64_t> shape;          // [ndim]
    std::vector<int64_t> stride[KMAXOPS];// [ndim] per operand, elements
    int64_t base[KMAXOPS] = {0, 0, 0};   // element offset of index 0 per operand
    int64_t total() const {
        int64_t s = 1;
        for (auto d : shape) s *= d;
        return s;
    }
};

// layouts[0] is the output.  `collapse_out_broadcast`: iteration order G (in-place unary / fill): axes
// where the output has stride 0 are visited once.  A broadcast output is otherwise rejected.
// `drop_common_broadcast`: an axis where the output AND every input have stride 0 is visited once instead of being
// rejected -- only valid when the output storage is not also an input (out-of-place ops).
CanonEw canon_elementwise(const std::vector<const Layout *> &layouts, bool collapse_out_broadcast,
                          bool drop_common_broadcast = false);

// Rewrites (lc, la) of equal SIZE but different shape into two layouts of one common refined shape such
// that same-index pairing equals flattened-order pairing in `order` (assign_arbitary semantics,
// cpu_serial/assignment.rs:39-67).  Returns false if the two shapes have no common refinement.
bool refine_to_common_shape(const Layout &lc, const Layout &la, rc_order order, Layout *oc, Layout *oa);

// Reduction descriptor: kept axes (input stride, output stride) and reduced axes (input stride).
struct CanonRed {
    bool empty_out = false;              // no output elements
    std::vector<int64_t> kshape, kstride_in, kstride_out;  // kept dims, dim 0 = fastest on the OUTPUT
    std::vector<int64_t> rshape, rstride;                  // reduced dims, dim 0 = smallest |stride|
    int64_t base_in = 0, base_out = 0;
    // second input of a binary reduction (vecdot, allclose): same dims, its own strides
    bool binary = false;
    std::vector<int64_t> kstride_in2, rstride2;
    int64_t base_in2 = 0;
    int64_t n_out() const { int64_t s = 1; for (auto d : kshape) s *= d; return s; }
    int64_t n_red() const { int64_t s = 1; for (auto d : rshape) s *= d; return s; }
};
// keep_order: the reduced dims keep the row-major order of `axes` (dim 0 = last axis given), are never flipped
// or sorted and only merged when adjacent -- the flat reduced index then IS the row-major index arg* ops return.
CanonRed canon_reduce(const Layout &la, const std::vector<int> &axes, const Layout &lo, bool keep_order = false);
// Binary reduction out[k] = fold_r f(a[k, r], b[k, r]): kept views lam / lbm already broadcast to lo's shape,
// reduced views las / lbs of one shape (offsets: la_offset / lb_offset, counted once).  The same flips, order and
// merges are applied to both inputs, so the element pairing is preserved.
CanonRed canon_reduce_binary(const Layout &lam, const Layout &lbm, const Layout &lo, const Layout &las, const Layout &lbs,
                             int64_t la_offset, int64_t lb_offset);

}  // namespace rc
