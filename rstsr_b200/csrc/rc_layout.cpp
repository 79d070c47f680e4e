// rc_layout.cpp -- host-side layout algebra of DeviceCuda.
//
// A drop-in device must hand back the SAME layouts the reference's L0/L4 code computes
// (SURVEY section 7 "hard parts": output strides of K-order ops depend on the input strides).
// Each function below states the reference function whose observable result it reproduces;
// the implementation is written from that behaviour, not transcribed.
#include "rc_layout.hpp"

#include <algorithm>
#include <numeric>

namespace rc {

static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
const std::string &last_error_ref() { return g_last_error; }

const char *dtype_name(rc_dtype t) {
    switch (t) {
        case RC_BOOL: return "bool";
        case RC_I8: return "i8";
        case RC_I16: return "i16";
        case RC_I32: return "i32";
        case RC_I64: return "i64";
        case RC_U8: return "u8";
        case RC_U16: return "u16";
        case RC_U32: return "u32";
        case RC_U64: return "u64";
        case RC_F32: return "f32";
        case RC_F64: return "f64";
        case RC_F16: return "f16";
        case RC_BF16: return "bf16";
        case RC_C32: return "c32";
        case RC_C64: return "c64";
    }
    return "?";
}

Layout from_c(const rc_layout *l) {
    RC_CHECK(l != nullptr, RC_ERR_INVALID_VALUE, "null layout");
    RC_CHECK(l->ndim >= 0 && l->ndim <= RC_MAX_NDIM, RC_ERR_INVALID_LAYOUT, "ndim out of range (0..16)");
    Layout r;
    r.shape.assign(l->shape, l->shape + l->ndim);
    r.stride.assign(l->stride, l->stride + l->ndim);
    r.offset = l->offset;
    for (auto d : r.shape) RC_CHECK(d >= 0, RC_ERR_INVALID_LAYOUT, "negative extent in shape");
    RC_CHECK(r.offset >= 0, RC_ERR_INVALID_LAYOUT, "negative offset");
    return r;
}

void to_c(const Layout &l, rc_layout *out) {
    RC_CHECK(out != nullptr, RC_ERR_INVALID_VALUE, "null output layout");
    RC_CHECK(l.ndim() <= RC_MAX_NDIM, RC_ERR_INVALID_LAYOUT, "ndim exceeds RC_MAX_NDIM");
    std::memset(out, 0, sizeof(*out));
    out->ndim = l.ndim();
    for (int i = 0; i < l.ndim(); ++i) {
        out->shape[i] = l.shape[i];
        out->stride[i] = l.stride[i];
    }
    out->offset = l.offset;
}

// Layout::bounds_index (rstsr-common/src/layout/layoutbase.rs:237-262)
void bounds_index(const Layout &l, int64_t *mn, int64_t *mx) {
    if (l.ndim() == 0) {
        *mn = l.offset;
        *mx = l.offset + 1;
        return;
    }
    int64_t lo = l.offset, hi = l.offset;
    for (int i = 0; i < l.ndim(); ++i) {
        if (l.shape[i] == 0) {
            *mn = *mx = l.offset;
            return;
        }
        int64_t span = l.stride[i] * (l.shape[i] - 1);
        if (l.stride[i] > 0) hi += span; else lo += span;
    }
    RC_CHECK(lo >= 0, RC_ERR_VALUE_OUT_OF_RANGE, "layout reaches below index 0");
    *mn = lo;
    *mx = hi + 1;
}

// Layout::check_strides (layoutbase.rs:285-320)
void check_strides(const Layout &l, bool skip_zero) {
    if (l.ndim() == 0 || l.size() == 0) return;
    std::vector<int> idx;
    for (int k = 0; k < l.ndim(); ++k)
        if (l.shape[k] > 1) idx.push_back(k);
    std::stable_sort(idx.begin(), idx.end(),
                     [&](int p, int q) { return std::llabs(l.stride[p]) < std::llabs(l.stride[q]); });
    int64_t covered = 0;  // furthest element reachable through the smaller-stride axes
    for (int k : idx) {
        int64_t t = std::llabs(l.stride[k]);
        if (t == 0 && skip_zero) continue;
        RC_CHECK(covered < t, RC_ERR_INVALID_LAYOUT,
                 "Either stride be zero, or stride too small that elements in tensor can be overlapped.");
        covered += (l.shape[k] - 1) * t;
    }
}

// Layout::new (layoutbase.rs:396-404)
void check_layout(const Layout &l) {
    RC_CHECK(l.shape.size() == l.stride.size(), RC_ERR_INVALID_LAYOUT, "shape/stride length mismatch");
    int64_t a, b;
    bounds_index(l, &a, &b);
    check_strides(l, true);
}

// Layout::ndim_of_f_contig / ndim_of_c_contig (layoutbase.rs:152-186)
int ndim_of_f_contig(const Layout &l) {
    if (l.ndim() == 0 || l.size() == 0) return l.ndim();
    int64_t expect = 1;
    for (int i = 0; i < l.ndim(); ++i) {
        if (l.shape[i] != 1 && l.stride[i] != expect) return i;
        expect *= l.shape[i];
    }
    return l.ndim();
}

int ndim_of_c_contig(const Layout &l) {
    if (l.ndim() == 0 || l.size() == 0) return l.ndim();
    int64_t expect = 1;
    for (int k = 0; k < l.ndim(); ++k) {
        int i = l.ndim() - 1 - k;
        if (l.shape[i] != 1 && l.stride[i] != expect) return k;
        expect *= l.shape[i];
    }
    return l.ndim();
}

// DimLayoutContigAPI::new_c_contig / new_f_contig (layoutbase.rs:574-601)
Layout new_contig(const std::vector<int64_t> &shape, rc_order order, int64_t offset) {
    Layout r;
    r.shape = shape;
    r.stride.assign(shape.size(), 0);
    r.offset = offset;
    int n = (int)shape.size();
    int64_t acc = 1;
    if (order == RC_ROW_MAJOR) {
        for (int i = n - 1; i >= 0; --i) {
            r.stride[i] = acc;
            acc *= std::max<int64_t>(shape[i], 1);
        }
    } else {
        for (int i = 0; i < n; ++i) {
            r.stride[i] = acc;
            acc *= std::max<int64_t>(shape[i], 1);
        }
    }
    return r;
}

Layout permuted(const Layout &l, const std::vector<int> &axes) {
    Layout r;
    r.offset = l.offset;
    r.shape.resize(axes.size());
    r.stride.resize(axes.size());
    for (size_t i = 0; i < axes.size(); ++i) {
        r.shape[i] = l.shape[axes[i]];
        r.stride[i] = l.stride[axes[i]];
    }
    return r;
}

Layout reversed_axes(const Layout &l) {
    Layout r = l;
    std::reverse(r.shape.begin(), r.shape.end());
    std::reverse(r.stride.begin(), r.stride.end());
    return r;
}

// PartialEq for Layout (layoutbase.rs:547-572): strides of extent-0/1 axes are irrelevant
bool layout_equal(const Layout &a, const Layout &b) {
    if (a.ndim() != b.ndim() || a.offset != b.offset) return false;
    for (int i = 0; i < a.ndim(); ++i) {
        if (a.shape[i] != b.shape[i]) return false;
        if (a.shape[i] > 1 && a.stride[i] != b.stride[i]) return false;
    }
    return true;
}

// Layout::size_non_broadcast (broadcast.rs:255-266)
int64_t size_non_broadcast(const Layout &l) {
    if (l.size() == 0) return 0;
    int64_t s = 1;
    for (int i = 0; i < l.ndim(); ++i)
        if (l.stride[i] != 0) s *= l.shape[i];
    return s;
}

// broadcast_layout (broadcast.rs:21-95,166-245).  Row-major aligns shapes at the right, col-major at
// the left; every axis that is missing or stretched from extent 1 gets stride 0.
static Layout stretch_row_major(const Layout &l, const std::vector<int64_t> &shape) {
    int n = (int)shape.size(), m = l.ndim();
    Layout r;
    r.shape = shape;
    r.stride.assign(n, 0);
    r.offset = l.offset;
    for (int i = 0; i < n; ++i) {
        int j = i - (n - m);  // matching axis of l
        if (j < 0) continue;  // expanded axis
        bool upcast = (l.shape[j] == 1 && shape[i] != 1);
        r.stride[i] = upcast ? 0 : l.stride[j];
    }
    return r;
}

void broadcast_layouts(const Layout &la_in, const Layout &lb_in, rc_order order, Layout *oa, Layout *ob) {
    Layout la = la_in, lb = lb_in;
    if (order == RC_COL_MAJOR) {
        la = reversed_axes(la);
        lb = reversed_axes(lb);
    }
    int na = la.ndim(), nb = lb.ndim(), n = std::max(na, nb);
    RC_CHECK(n <= RC_MAX_NDIM, RC_ERR_INVALID_LAYOUT, "broadcast ndim exceeds RC_MAX_NDIM");
    std::vector<int64_t> shape(n, 1);
    for (int i = 0; i < n; ++i) {
        int ja = i - (n - na), jb = i - (n - nb);
        int64_t da = ja >= 0 ? la.shape[ja] : 1;
        int64_t db = jb >= 0 ? lb.shape[jb] : 1;
        if (da == 1) shape[i] = db;
        else if (db == 1) shape[i] = da;
        else {
            RC_CHECK(da == db, RC_ERR_INVALID_LAYOUT, "Broadcasting failed.");
            shape[i] = da;
        }
    }
    Layout ra = stretch_row_major(la, shape), rb = stretch_row_major(lb, shape);
    if (order == RC_COL_MAJOR) {
        ra = reversed_axes(ra);
        rb = reversed_axes(rb);
    }
    *oa = ra;
    *ob = rb;
}

// get_axes_composition (rearrangement.rs:335-372)
AxesComposition axes_composition(const Layout &l) {
    AxesComposition c;
    std::vector<int> rest;
    for (int i = 0; i < l.ndim(); ++i) {
        if (l.shape[i] == 1) c.one.push_back(i);
        else if (l.stride[i] == 0) c.zero.push_back(i);
        else rest.push_back(i);
    }
    std::stable_sort(rest.begin(), rest.end(),
                     [&](int p, int q) { return std::llabs(l.stride[p]) < std::llabs(l.stride[q]); });
    int64_t expect = 1;
    for (int i : rest) {
        if (l.stride[i] == expect) {
            c.contig.push_back(i);
            expect *= l.shape[i];
        } else {
            c.discontig.push_back(i);
        }
    }
    return c;
}

static bool contains(const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// get_layout_for_binary_op (rearrangement.rs:394-459)
Layout layout_for_binary_op(const Layout &la, const Layout &lb, rc_order order) {
    RC_CHECK(la.shape == lb.shape, RC_ERR_INVALID_LAYOUT, "Shape of two layouts must be the same for this function.");
    int n = la.ndim();
    AxesComposition ca = axes_composition(la), cb = axes_composition(lb);
    std::vector<int> zero_both;
    for (int i : ca.zero)
        if (contains(cb.zero, i)) zero_both.push_back(i);
    std::vector<int> lead;  // common prefix of both operands' contiguous axes
    for (size_t k = 0; k < std::min(ca.contig.size(), cb.contig.size()); ++k) {
        if (ca.contig[k] != cb.contig[k]) break;
        lead.push_back(ca.contig[k]);
    }
    std::vector<int> tail;
    for (int i = 0; i < n; ++i)
        if (!contains(zero_both, i) && !contains(lead, i) && !contains(ca.one, i)) tail.push_back(i);
    if (order == RC_ROW_MAJOR) std::reverse(tail.begin(), tail.end());

    Layout lc;
    lc.shape = la.shape;
    lc.stride.assign(n, 0);
    lc.offset = 0;
    int64_t acc = 1;
    for (int i : lead) { lc.stride[i] = acc; acc *= la.shape[i]; }
    for (int i : tail) { lc.stride[i] = acc; acc *= la.shape[i]; }
    // extent-1 axes borrow the nearest non-zero stride (to the right for row-major, left for col-major)
    for (int i : ca.one) {
        int64_t s = 1;
        if (order == RC_ROW_MAJOR) {
            for (int j = i; j < n; ++j)
                if (lc.stride[j] != 0) { s = lc.stride[j]; break; }
        } else {
            for (int j = i - 1; j >= 0; --j)
                if (lc.stride[j] != 0) { s = lc.stride[j]; break; }
        }
        lc.stride[i] = s;
    }
    return lc;
}

// greedy_layout (rearrangement.rs:36-113).  Stable ordering: non-moving axes (extent 1 / stride 0) first
// (keep_shape) or last (!keep_shape) in index order; the others by ascending |stride|.
Layout greedy_layout(const Layout &l_in, bool keep_shape, std::vector<int> *perm_out) {
    Layout l = l_in;
    int n = l.ndim();
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    if (l.size() == 0) {
        if (perm_out) *perm_out = perm;
        return l;
    }
    if (keep_shape) {
        for (int i = 0; i < n; ++i)
            if (l.stride[i] < 0) {  // view the axis backwards: same elements, positive stride
                l.offset += (l.shape[i] - 1) * l.stride[i];
                l.stride[i] = -l.stride[i];
            }
    }
    auto still = [&](int i) { return l.shape[i] == 1 || l.stride[i] == 0; };
    std::stable_sort(perm.begin(), perm.end(), [&](int p, int q) {
        bool sp = still(p), sq = still(q);
        if (sp && sq) return p < q;
        if (sp != sq) return keep_shape ? sp : sq;
        return std::llabs(l.stride[p]) < std::llabs(l.stride[q]);
    });
    Layout g = permuted(l, perm);
    if (!keep_shape) {
        for (int i = 0; i < n; ++i)
            if (g.shape[i] == 1 || g.stride[i] == 0) { g.shape[i] = 1; g.stride[i] = 0; }
    }
    if (perm_out) *perm_out = perm;
    return g;
}

// layout_for_array_copy (rearrangement.rs:125-152)
Layout layout_for_array_copy(const Layout &l, rc_iter_order it, rc_order default_order) {
    switch (it) {
        case RC_ITER_C: return new_contig(l.shape, RC_ROW_MAJOR, 0);
        case RC_ITER_F: return new_contig(l.shape, RC_COL_MAJOR, 0);
        case RC_ITER_A:
            if (c_contig(l)) return new_contig(l.shape, RC_ROW_MAJOR, 0);
            if (f_contig(l)) return new_contig(l.shape, RC_COL_MAJOR, 0);
            return new_contig(l.shape, default_order, 0);
        case RC_ITER_K: {
            std::vector<int> perm;
            Layout g = greedy_layout(l, true, &perm);
            Layout f = new_contig(g.shape, RC_COL_MAJOR, 0);
            std::vector<int> inv(perm.size());
            for (size_t k = 0; k < perm.size(); ++k) inv[perm[k]] = (int)k;
            return permuted(f, inv);
        }
    }
    raise(RC_ERR_INVALID_VALUE, "Iter order for copy only accepts CFAK.");
}

// normalize_axes_index (rstsr-common/src/axis_index.rs:379-414), allow_duplicate = false, sort = false
std::vector<int> normalize_axes(const int64_t *axes, int naxes, int ndim) {
    RC_CHECK(naxes >= 0 && (naxes == 0 || axes != nullptr), RC_ERR_INVALID_VALUE, "invalid axes argument");
    std::vector<int> r;
    for (int k = 0; k < naxes; ++k) {
        int64_t a = axes[k];
        if (a < 0) a += ndim;
        RC_CHECK(a >= 0 && a < ndim, RC_ERR_INVALID_VALUE, "axis out of bounds for the number of dimensions");
        r.push_back((int)a);
    }
    std::vector<int> s = r;
    std::sort(s.begin(), s.end());
    for (size_t k = 1; k < s.size(); ++k)
        RC_CHECK(s[k] != s[k - 1], RC_ERR_INVALID_VALUE, "Duplicate axes are not allowed.");
    return r;
}

// Layout::dim_split_axes (rstsr-common/src/layout/indexer.rs:453-478); both halves keep the offset
void split_axes(const Layout &l, const std::vector<int> &axes, Layout *l_axes, Layout *l_rest,
                std::vector<int> *rest_axes) {
    std::vector<int> rest;
    for (int i = 0; i < l.ndim(); ++i)
        if (!contains(axes, i)) rest.push_back(i);
    Layout a = permuted(l, axes), r = permuted(l, rest);
    check_layout(a);
    check_layout(r);
    if (l_axes) *l_axes = a;
    if (l_rest) *l_rest = r;
    if (rest_axes) *rest_axes = rest;
}

// output layout of reduce_axes (cpu_rayon/reduction.rs:147-153)
Layout layout_for_reduce(const Layout &la, const std::vector<int> &axes) {
    Layout kept;
    split_axes(la, axes, nullptr, &kept, nullptr);
    return layout_for_array_copy(kept, RC_ITER_K, RC_ROW_MAJOR);
}

// layout_reshapeable (rstsr-common/src/layout/reshape.rs:8-226).  The no-copy test is NumPy's
// _attempt_nocopy_reshape: walk both shapes, grouping axes whose extents multiply to the same count;
// a group can be viewed iff the old axes in it are mutually contiguous in the requested order.
static bool nocopy_strides(const Layout &la, const std::vector<int64_t> &nshape, bool f_order,
                           std::vector<int64_t> *out) {
    std::vector<int64_t> od, os;
    for (int i = 0; i < la.ndim(); ++i)
        if (la.shape[i] != 1) { od.push_back(la.shape[i]); os.push_back(la.stride[i]); }
    int on = (int)od.size(), nn = (int)nshape.size();
    std::vector<int64_t> ns(nn, 0);
    int oi = 0, oj = 1, ni = 0, nj = 1;
    while (ni < nn && oi < on) {
        int64_t np = nshape[ni], op = od[oi];
        while (np != op) {
            if (np < op) {
                if (nj >= nn) return false;
                np *= nshape[nj++];
            } else {
                if (oj >= on) return false;
                op *= od[oj++];
            }
        }
        for (int k = oi; k + 1 < oj; ++k) {
            if (f_order) { if (os[k + 1] != od[k] * os[k]) return false; }
            else         { if (os[k] != od[k + 1] * os[k + 1]) return false; }
        }
        if (f_order) {
            ns[ni] = os[oi];
            for (int k = ni + 1; k < nj; ++k) ns[k] = ns[k - 1] * nshape[k - 1];
        } else {
            ns[nj - 1] = os[oj - 1];
            for (int k = nj - 1; k > ni; --k) ns[k - 1] = ns[k] * nshape[k];
        }
        ni = nj++;
        oi = oj++;
    }
    int64_t last = 1;
    if (ni >= 1) {
        last = ns[ni - 1];
        if (f_order) last *= nshape[ni - 1];
    }
    for (int k = ni; k < nn; ++k) ns[k] = last;
    *out = ns;
    return true;
}

bool reshapeable(const Layout &la, const std::vector<int64_t> &shape, rc_order order, Layout *out) {
    int64_t size_out = 1;
    for (auto d : shape) {
        RC_CHECK(d >= 0, RC_ERR_INVALID_VALUE, "negative extent in target shape");
        if (d != 0 && size_out > INT64_MAX / d) raise(RC_ERR_INVALID_VALUE, "Output shape product overflows.");
        size_out *= d;
    }
    RC_CHECK(size_out == la.size(), RC_ERR_INVALID_VALUE, "Size mismatch between input tensor and output tensor.");
    if (size_out == 0 || size_out == 1) {
        Layout r;
        r.shape = shape;
        r.stride.assign(shape.size(), 1);
        r.offset = la.offset;
        check_layout(r);
        *out = r;
        return true;
    }
    if (shape == la.shape) { *out = la; return true; }
    if (order == RC_ROW_MAJOR ? c_contig(la) : f_contig(la)) {
        *out = new_contig(shape, order, la.offset);
        return true;
    }
    std::vector<int64_t> ns;
    if (!nocopy_strides(la, shape, order == RC_COL_MAJOR, &ns)) return false;
    out->shape = shape;
    out->stride = ns;
    out->offset = la.offset;
    return true;
}

}  // namespace rc
