// rc_layout.hpp -- host-side layout algebra that must agree with the reference bit for bit
// (which axes broadcast, which strides an output gets).  See rc_layout.cpp for citations.
#pragma once
#include "rc_common.hpp"

namespace rc {

// [min, max) element-index bounds of a layout; error if min < 0.
void bounds_index(const Layout &l, int64_t *mn, int64_t *mx);
// overlap check used by Layout::new (stride-0 axes skipped)
void check_strides(const Layout &l, bool skip_zero);
void check_layout(const Layout &l);

int ndim_of_f_contig(const Layout &l);
int ndim_of_c_contig(const Layout &l);
inline bool f_contig(const Layout &l) { return ndim_of_f_contig(l) == l.ndim(); }
inline bool c_contig(const Layout &l) { return ndim_of_c_contig(l) == l.ndim(); }

Layout new_contig(const std::vector<int64_t> &shape, rc_order order, int64_t offset);
Layout permuted(const Layout &l, const std::vector<int> &axes);
Layout reversed_axes(const Layout &l);
bool layout_equal(const Layout &a, const Layout &b);
int64_t size_non_broadcast(const Layout &l);

void broadcast_layouts(const Layout &la, const Layout &lb, rc_order order, Layout *oa, Layout *ob);

// axes grouped as (size-1, stride-0, memory-contiguous ascending, remaining ascending by |stride|)
struct AxesComposition {
    std::vector<int> one, zero, contig, discontig;
};
AxesComposition axes_composition(const Layout &l);

Layout layout_for_binary_op(const Layout &la, const Layout &lb, rc_order order);

// "greedy" (memory-following) axis order; keep_shape = true is iteration order K, false is order G.
Layout greedy_layout(const Layout &l, bool keep_shape, std::vector<int> *perm);
Layout layout_for_array_copy(const Layout &l, rc_iter_order it, rc_order default_order);

std::vector<int> normalize_axes(const int64_t *axes, int naxes, int ndim);
// (layout of `axes` in the given order, layout of the remaining axes in ascending order)
void split_axes(const Layout &l, const std::vector<int> &axes, Layout *l_axes, Layout *l_rest,
                std::vector<int> *rest_axes);
Layout layout_for_reduce(const Layout &la, const std::vector<int> &axes);

bool reshapeable(const Layout &la, const std::vector<int64_t> &shape, rc_order order, Layout *out);

}  // namespace rc
