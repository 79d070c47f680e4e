// rc_canon.cpp -- host canonicaliser (see rc_canon.hpp).
#include "rc_canon.hpp"

#include <algorithm>
#include <numeric>

#include "rc_layout.hpp"

namespace rc {

namespace {
struct Dim {
    int64_t n;
    int64_t s[KMAXOPS];
};
}  // namespace

CanonEw canon_elementwise(const std::vector<const Layout *> &ls, bool collapse_out_broadcast, bool drop_common_broadcast) {
    CanonEw c;
    c.nops = (int)ls.size();
    RC_CHECK(c.nops >= 1 && c.nops <= KMAXOPS, RC_ERR_RUNTIME, "operand count");
    const Layout &lo = *ls[0];
    for (int k = 1; k < c.nops; ++k)
        RC_CHECK(ls[k]->shape == lo.shape, RC_ERR_INVALID_LAYOUT,
                 "All shape of layout in this function must be the same.");
    for (int k = 0; k < c.nops; ++k) {
        RC_CHECK(ls[k]->stride.size() == ls[k]->shape.size(), RC_ERR_INVALID_LAYOUT, "shape/stride length mismatch");
        c.base[k] = ls[k]->offset;
    }
    std::vector<Dim> dims;
    for (int i = 0; i < lo.ndim(); ++i) {
        if (lo.shape[i] == 0) {
            c.empty = true;
            return c;
        }
        if (lo.shape[i] == 1) continue;
        Dim d{};
        d.n = lo.shape[i];
        for (int k = 0; k < c.nops; ++k) d.s[k] = ls[k]->stride[i];
        if (d.s[0] == 0) {
            if (collapse_out_broadcast) continue;  // order G: a broadcast output axis is visited once
            // both inputs broadcast along the axis too (x.broadcast_to(s) + y.broadcast_to(s): get_layout_for_binary_op
            // gives the output stride 0 there): every visit would write the same value, so the axis is visited once
            bool all_zero = drop_common_broadcast;
            for (int k = 1; k < c.nops && all_zero; ++k) all_zero = d.s[k] == 0;
            if (all_zero) continue;
            raise(RC_ERR_INVALID_LAYOUT, "output layout is broadcast (stride 0 on an axis of extent > 1)");
        }
        if (d.s[0] < 0) {  // walk the axis the other way round in every operand
            for (int k = 0; k < c.nops; ++k) {
                c.base[k] += (d.n - 1) * d.s[k];
                d.s[k] = -d.s[k];
            }
        }
        dims.push_back(d);
    }
    std::stable_sort(dims.begin(), dims.end(), [](const Dim &a, const Dim &b) { return a.s[0] < b.s[0]; });
    // merge dim i+1 into dim i when every operand steps contiguously from one to the other
    std::vector<Dim> m;
    for (const Dim &d : dims) {
        if (!m.empty()) {
            Dim &p = m.back();
            bool ok = true;
            for (int k = 0; k < c.nops; ++k) ok = ok && (d.s[k] == p.s[k] * p.n);
            if (ok) {
                p.n *= d.n;
                continue;
            }
        }
        m.push_back(d);
    }
    if (m.empty()) {
        Dim d;
        d.n = 1;
        for (int k = 0; k < KMAXOPS; ++k) d.s[k] = 0;
        m.push_back(d);
    }
    c.ndim = (int)m.size();
    for (const Dim &d : m) {
        c.shape.push_back(d.n);
        for (int k = 0; k < c.nops; ++k) c.stride[k].push_back(d.s[k]);
    }
    return c;
}

bool refine_to_common_shape(const Layout &lc_in, const Layout &la_in, rc_order order, Layout *oc, Layout *oa) {
    // Work in "fastest axis first" order: reversed for row-major, as is for col-major.
    Layout lc = lc_in, la = la_in;
    if (order == RC_ROW_MAJOR) {
        lc = reversed_axes(lc);
        la = reversed_axes(la);
    }
    // drop extent-1 axes and fuse neighbours that are contiguous with each other, so that e.g. a
    // C-contiguous target always becomes one flat run (which refines against every shape)
    auto strip = [](const Layout &l, std::vector<int64_t> *n, std::vector<int64_t> *s) {
        for (int i = 0; i < l.ndim(); ++i) {
            if (l.shape[i] == 1) continue;
            if (!n->empty() && l.stride[i] == s->back() * n->back()) { n->back() *= l.shape[i]; continue; }
            n->push_back(l.shape[i]);
            s->push_back(l.stride[i]);
        }
    };
    std::vector<int64_t> nc, sc, na, sa;
    strip(lc, &nc, &sc);
    strip(la, &na, &sa);
    std::vector<int64_t> shape, rc_s, ra_s;
    size_t ic = 0, ia = 0;
    int64_t remc = nc.empty() ? 1 : nc[0], rema = na.empty() ? 1 : na[0];
    int64_t curc = sc.empty() ? 0 : sc[0], cura = sa.empty() ? 0 : sa[0];
    while (ic < nc.size() && ia < na.size()) {
        int64_t take = std::min(remc, rema);
        if (remc % take != 0 || rema % take != 0) return false;
        shape.push_back(take);
        rc_s.push_back(curc);
        ra_s.push_back(cura);
        remc /= take; curc *= take;
        rema /= take; cura *= take;
        if (remc == 1) { if (++ic < nc.size()) { remc = nc[ic]; curc = sc[ic]; } }
        if (rema == 1) { if (++ia < na.size()) { rema = na[ia]; cura = sa[ia]; } }
    }
    if (ic < nc.size() || ia < na.size()) return false;  // sizes differ (caller checks) or leftover
    Layout c, a;
    c.shape = shape; c.stride = rc_s; c.offset = lc_in.offset;
    a.shape = shape; a.stride = ra_s; a.offset = la_in.offset;
    *oc = c;
    *oa = a;
    return true;
}

CanonRed canon_reduce(const Layout &la, const std::vector<int> &axes, const Layout &lo, bool keep_order) {
    CanonRed c;
    c.base_in = la.offset;
    c.base_out = lo.offset;
    std::vector<int> kept;
    for (int i = 0; i < la.ndim(); ++i)
        if (std::find(axes.begin(), axes.end(), i) == axes.end()) kept.push_back(i);
    RC_CHECK((int)kept.size() == lo.ndim(), RC_ERR_INVALID_LAYOUT, "output layout rank must equal the kept axes");
    struct K { int64_t n, si, so; };
    std::vector<K> ks;
    for (size_t j = 0; j < kept.size(); ++j) {
        int i = kept[j];
        RC_CHECK(lo.shape[j] == la.shape[i], RC_ERR_INVALID_LAYOUT, "output shape must equal the kept axes' shape");
        if (la.shape[i] == 0) { c.empty_out = true; return c; }
        if (la.shape[i] == 1) continue;
        K k{la.shape[i], la.stride[i], lo.stride[j]};
        RC_CHECK(k.so != 0, RC_ERR_INVALID_LAYOUT, "output layout is broadcast (stride 0 on an axis of extent > 1)");
        if (k.so < 0) {
            c.base_in += (k.n - 1) * k.si; k.si = -k.si;
            c.base_out += (k.n - 1) * k.so; k.so = -k.so;
        }
        ks.push_back(k);
    }
    std::stable_sort(ks.begin(), ks.end(), [](const K &a, const K &b) { return a.so < b.so; });
    for (const K &k : ks) {
        if (!c.kshape.empty()) {
            size_t p = c.kshape.size() - 1;
            if (k.si == c.kstride_in[p] * c.kshape[p] && k.so == c.kstride_out[p] * c.kshape[p]) {
                c.kshape[p] *= k.n;
                continue;
            }
        }
        c.kshape.push_back(k.n);
        c.kstride_in.push_back(k.si);
        c.kstride_out.push_back(k.so);
    }
    struct R { int64_t n, s; };
    std::vector<R> rs;
    for (size_t k = 0; k < axes.size(); ++k) {
        // keep_order: fastest reduced dim = LAST axis given (row-major flattening of the reduced space)
        int i = keep_order ? axes[axes.size() - 1 - k] : axes[k];
        if (la.shape[i] == 1) continue;
        R r{la.shape[i], la.stride[i]};
        if (!keep_order && r.s < 0 && r.n > 0) {  // an order-free reduction may walk any axis in either direction
            c.base_in += (r.n - 1) * r.s;
            r.s = -r.s;
        }
        rs.push_back(r);
    }
    if (!keep_order) std::stable_sort(rs.begin(), rs.end(), [](const R &a, const R &b) { return a.s < b.s; });
    for (const R &r : rs) {
        if (!c.rshape.empty()) {
            size_t p = c.rshape.size() - 1;
            if (r.s == c.rstride[p] * c.rshape[p]) {
                c.rshape[p] *= r.n;
                continue;
            }
        }
        c.rshape.push_back(r.n);
        c.rstride.push_back(r.s);
    }
    return c;
}

CanonRed canon_reduce_binary(const Layout &lam, const Layout &lbm, const Layout &lo, const Layout &las, const Layout &lbs,
                             int64_t la_offset, int64_t lb_offset) {
    CanonRed c;
    c.binary = true;
    c.base_in = la_offset;
    c.base_in2 = lb_offset;
    c.base_out = lo.offset;
    RC_CHECK(lam.shape == lo.shape && lbm.shape == lo.shape, RC_ERR_INVALID_LAYOUT,
             "kept axes of both inputs must be broadcast to the output shape");
    RC_CHECK(las.shape == lbs.shape, RC_ERR_INVALID_LAYOUT, "reduced axes of both inputs must have one shape");
    struct K { int64_t n, sa, sb, so; };
    std::vector<K> ks;
    for (int j = 0; j < lo.ndim(); ++j) {
        if (lo.shape[j] == 0) { c.empty_out = true; return c; }
        if (lo.shape[j] == 1) continue;
        K k{lo.shape[j], lam.stride[j], lbm.stride[j], lo.stride[j]};
        if (k.so == 0 && k.sa == 0 && k.sb == 0) continue;  // kept axis broadcast in both inputs and the output: once
        RC_CHECK(k.so != 0, RC_ERR_INVALID_LAYOUT, "output layout is broadcast (stride 0 on an axis of extent > 1)");
        if (k.so < 0) {
            c.base_in += (k.n - 1) * k.sa; k.sa = -k.sa;
            c.base_in2 += (k.n - 1) * k.sb; k.sb = -k.sb;
            c.base_out += (k.n - 1) * k.so; k.so = -k.so;
        }
        ks.push_back(k);
    }
    std::stable_sort(ks.begin(), ks.end(), [](const K &x, const K &y) { return x.so < y.so; });
    for (const K &k : ks) {
        if (!c.kshape.empty()) {
            size_t p = c.kshape.size() - 1;
            if (k.sa == c.kstride_in[p] * c.kshape[p] && k.sb == c.kstride_in2[p] * c.kshape[p] &&
                k.so == c.kstride_out[p] * c.kshape[p]) {
                c.kshape[p] *= k.n;
                continue;
            }
        }
        c.kshape.push_back(k.n);
        c.kstride_in.push_back(k.sa);
        c.kstride_in2.push_back(k.sb);
        c.kstride_out.push_back(k.so);
    }
    struct R { int64_t n, sa, sb; };
    std::vector<R> rs;
    for (int i = 0; i < las.ndim(); ++i) {
        if (las.shape[i] == 1) continue;
        R r{las.shape[i], las.stride[i], lbs.stride[i]};
        if (r.n > 0 && (r.sa < 0 || (r.sa == 0 && r.sb < 0))) {  // walk the pair backwards together
            c.base_in += (r.n - 1) * r.sa; r.sa = -r.sa;
            c.base_in2 += (r.n - 1) * r.sb; r.sb = -r.sb;
        }
        rs.push_back(r);
    }
    auto key = [](const R &r) { return r.sa != 0 ? r.sa : (r.sb < 0 ? -r.sb : r.sb); };
    std::stable_sort(rs.begin(), rs.end(), [&](const R &x, const R &y) { return key(x) < key(y); });
    for (const R &r : rs) {
        if (!c.rshape.empty()) {
            size_t p = c.rshape.size() - 1;
            if (r.sa == c.rstride[p] * c.rshape[p] && r.sb == c.rstride2[p] * c.rshape[p]) {
                c.rshape[p] *= r.n;
                continue;
            }
        }
        c.rshape.push_back(r.n);
        c.rstride.push_back(r.sa);
        c.rstride2.push_back(r.sb);
    }
    return c;
}

}  // namespace rc
