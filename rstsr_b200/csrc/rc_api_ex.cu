// rc_api_ex.cu -- boundary completions on top of the core entry points:
//   * mixed operand types of the binary-function traits (promotion as rstsr-dtype-traits/src/promotion.rs:186-300,
//     rules of rstsr-core/src/feature_rayon/auto_impl/op_ternary_common.rs:6-189),
//   * elementwise isclose (OpIsCloseAPI, rstsr-core/src/operators/ops/op_ternary_common.rs:59-102),
//   * NUMA-placed pinned staging buffers (host side of outof_cpu_vec / to_cpu_vec on a two-socket box).
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <fstream>
#include <map>

#include "rc_canon.hpp"
#include "rc_device.hpp"
#include "rc_elementwise.cuh"
#include "rc_functors.cuh"
#include "rc_layout.hpp"
#include "rc_ops.hpp"

namespace rc {
namespace {

// ---- DTypePromoteAPI<TB> for TA (NumPy's table; usize = u64, isize = i64) ----
int int_bits(rc_dtype t) {
    switch (t) {
        case RC_I8: case RC_U8: return 8;
        case RC_I16: case RC_U16: return 16;
        case RC_I32: case RC_U32: return 32;
        case RC_I64: case RC_U64: return 64;
        default: return 0;
    }
}
rc_dtype signed_of_bits(int bits) { return bits <= 8 ? RC_I8 : bits <= 16 ? RC_I16 : bits <= 32 ? RC_I32 : RC_I64; }
rc_dtype unsigned_of_bits(int bits) { return bits <= 8 ? RC_U8 : bits <= 16 ? RC_U16 : bits <= 32 ? RC_U32 : RC_U64; }

rc_dtype promote(rc_dtype a, rc_dtype b) {
    if (a == b) return a;
    if (a == RC_BOOL) return b;  // bool x T -> T (promotion.rs:123-181; f16 / bf16 / c32 / c64: :195-200)
    if (b == RC_BOOL) return a;
    if (dtype_is_extended(a) || dtype_is_extended(b)) {
        const bool ca = dtype_is_complex(a), cb = dtype_is_complex(b);
        if (ca && cb) return RC_C64;  // c32 x c64 (promotion.rs:516-545)
        if ((ca || cb) && !dtype_is_half(a) && !dtype_is_half(b)) {
            const rc_dtype c = ca ? a : b, p = ca ? b : a;
            if (c == RC_C64) return RC_C64;  // Complex<f64> x primitive (:412-423, :493-505)
            // Complex<f32> keeps f32 components for the primitives f32 holds exactly (:406-410), else c64 (:425-431)
            return (p == RC_I8 || p == RC_I16 || p == RC_U8 || p == RC_U16 || p == RC_F32) ? RC_C32 : RC_C64;
        }
        raise(RC_ERR_UNIMPLEMENTED, std::string("DTypePromoteAPI<") + dtype_name(b) + "> is not implemented for " + dtype_name(a) +
                                        " (the half types pair with themselves and bool only, as in the reference)");
    }
    const bool fa = dtype_is_float(a), fb = dtype_is_float(b);
    if (fa && fb) return RC_F64;  // f32 x f64
    if (fa || fb) {
        const rc_dtype f = fa ? a : b, i = fa ? b : a;
        // f32 holds 8- and 16-bit integers exactly; wider integers force f64
        if (f == RC_F32) return int_bits(i) <= 16 ? RC_F32 : RC_F64;
        return RC_F64;
    }
    const bool sa = dtype_is_signed_int(a), sb = dtype_is_signed_int(b);
    const int ba = int_bits(a), bb = int_bits(b);
    if (sa == sb) return sa ? signed_of_bits(std::max(ba, bb)) : unsigned_of_bits(std::max(ba, bb));
    // signed x unsigned: the signed type if it is strictly wider, else the next wider signed type; u64 -> f64
    const int bs = sa ? ba : bb, bu = sa ? bb : ba;
    if (bs > bu) return signed_of_bits(bs);
    if (bu >= 64) return RC_F64;
    return signed_of_bits(bu * 2);
}

rc_dtype into_float(rc_dtype t) {  // DTypeIntoFloatAPI (promotion.rs:62-118): integers -> f64
    if (t == RC_F32 || t == RC_F64 || dtype_is_extended(t)) return t;  // promotion.rs:85-101
    RC_CHECK(t != RC_BOOL, RC_ERR_UNIMPLEMENTED, "bool has no float type (DTypeIntoFloatAPI is not implemented for bool)");
    return RC_F64;
}

bool is_cmp_op(rc_binop op) { return op >= RC_EQ && op <= RC_GE; }
bool is_float_func(rc_binop op) {
    return op == RC_ATAN2 || op == RC_COPYSIGN || op == RC_HYPOT || op == RC_LOGADDEXP || op == RC_NEXTAFTER;
}

enum PowKind { POW_SAME = 0, POW_FLOAT_INT = 1, POW_INT_UINT = 2 };
PowKind pow_kind(rc_dtype ta, rc_dtype tb) {
    if (dtype_is_half(ta) && tb == ta) return POW_SAME;  // num_traits::Float::powf of the half types (through f32)
    if (dtype_is_float(ta)) {
        if (tb == ta) return POW_SAME;
        if (tb == RC_I8 || tb == RC_U8 || tb == RC_I16 || tb == RC_U16 || tb == RC_I32) return POW_FLOAT_INT;
    } else if (dtype_is_int(ta)) {
        if (dtype_is_unsigned_int(tb)) return POW_INT_UINT;
    }
    raise(RC_ERR_UNIMPLEMENTED, std::string("pow is not implemented for ") + dtype_name(ta) + " ^ " + dtype_name(tb) +
                                    " (num::Pow has no such impl)");
}

// compute type K (both operands are brought to it) and output type of `op` for operand types (ta, tb)
void op_types(rc_binop op, rc_dtype ta, rc_dtype tb, rc_dtype *k, rc_dtype *out) {
    if (dtype_is_extended(ta) || dtype_is_extended(tb)) {
        // half / complex operands: both are brought to promote(ta, tb) (bool x T, complex x primitive, c32 x c64 -- the rows
        // of the reference's table; anything else raises), then the one-type kernel of that dtype runs
        if (op == RC_POW && ta != tb)
            raise(RC_ERR_UNIMPLEMENTED, std::string("pow is not implemented for ") + dtype_name(ta) + " ^ " + dtype_name(tb));
        const rc_dtype r = promote(ta, tb);
        *k = r;
        *out = is_cmp_op(op) ? RC_BOOL : r;
        return;
    }
    if (op == RC_POW) {
        pow_kind(ta, tb);
        *k = ta;
        *out = ta;
        return;
    }
    const rc_dtype r = promote(ta, tb);
    if (is_float_func(op)) { *k = into_float(r); *out = *k; return; }
    *k = r;
    *out = is_cmp_op(op) ? RC_BOOL : r;
}

// An operand brought to dtype `want`: itself, or a compact cast copy (broadcast axes keep extent 1 in the copy and
// stride 0 in the view), freed stream-ordered when the holder goes out of scope.
struct Operand {
    rc_device *dev = nullptr;
    const void *ptr = nullptr;
    Layout l;
    void *tmp = nullptr;
    ~Operand() { if (tmp) cudaFreeAsync(tmp, dev->stream); }
};

void prepare(rc_device *dev, rc_dtype want, rc_dtype have, const void *ptr, const Layout &l, Operand *o) {
    o->dev = dev;
    if (want == have) { o->ptr = ptr; o->l = l; return; }
    Layout small = l;
    for (int i = 0; i < l.ndim(); ++i)
        if (l.stride[i] == 0) small.shape[i] = 1;
    const int64_t n = small.size();
    Layout tl = new_contig(small.shape, RC_ROW_MAJOR, 0);
    if (n > 0) {
        cudaError_t e = cudaMallocAsync(&o->tmp, (size_t)n * dtype_size(want), dev->stream);
        if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
        rc_layout c_tl, c_small;
        to_c(tl, &c_tl);
        to_c(small, &c_small);
        int st = rc_assign(dev, want, o->tmp, &c_tl, have, ptr, &c_small);
        if (st != RC_OK) raise((rc_status)st, rc_last_error());
    }
    o->ptr = o->tmp ? o->tmp : ptr;
    o->l.shape = l.shape;
    o->l.stride.assign(l.ndim(), 0);
    for (int i = 0; i < l.ndim(); ++i) o->l.stride[i] = (l.stride[i] == 0) ? 0 : tl.stride[i];
    o->l.offset = 0;
}

void check_out_type(rc_binop op, rc_dtype tc, rc_dtype ta, rc_dtype tb, rc_dtype *k) {
    rc_dtype out;
    op_types(op, ta, tb, k, &out);
    RC_CHECK(tc == out, RC_ERR_INVALID_VALUE,
             std::string("output dtype must be ") + dtype_name(out) + " for these operand types (got " + dtype_name(tc) + ")");
}

void status(int st) {
    if (st != RC_OK) raise((rc_status)st, rc_last_error());
}

// flat row-major index within `shape` -> index tuple, one thread per output element
struct UnravelDesc {
    int n;
    int64_t shape[RC_MAX_NDIM];
};
__global__ void unravel_kernel(const __grid_constant__ UnravelDesc d, const uint64_t *__restrict__ flat,
                               uint64_t *__restrict__ out, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t idx = flat[i];
    for (int k = d.n - 1; k >= 0; --k) {
        const uint64_t e = (uint64_t)d.shape[k];
        const uint64_t q = idx / e;
        out[i * d.n + k] = idx - q * e;
        idx = q;
    }
}

// ---- NUMA-bound pinned buffers ----
std::mutex g_numa_mu;
std::map<void *, size_t> g_numa_blocks;  // mmap'ed + cudaHostRegister'ed

constexpr int MPOL_BIND_ = 2;

// CPUs of a NUMA node that this process may run on (/sys/devices/system/node/nodeN/cpulist, e.g. "0-31,64-95")
bool node_cpus(int node, cpu_set_t *out) {
    std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string text;
    if (!f || !std::getline(f, text)) return false;
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return false;
    int n = 0;
    size_t i = 0;
    while (i < text.size()) {
        size_t j = text.find(',', i);
        std::string part = text.substr(i, j == std::string::npos ? std::string::npos : j - i);
        size_t dash = part.find('-');
        int lo = atoi(part.c_str()), hi = dash == std::string::npos ? lo : atoi(part.c_str() + dash + 1);
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) { CPU_SET(c, out); ++n; }
        if (j == std::string::npos) break;
        i = j + 1;
    }
    return n > 0;
}

}  // namespace

bool host_free_numa(void *ptr) {
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lock(g_numa_mu);
        auto it = g_numa_blocks.find(ptr);
        if (it == g_numa_blocks.end()) return false;
        bytes = it->second;
        g_numa_blocks.erase(it);
    }
    cudaHostUnregister(ptr);
    munmap(ptr, bytes);
    return true;
}

}  // namespace rc

using namespace rc;

extern "C" {

int rc_dtype_promote(rc_dtype ta, rc_dtype tb, rc_dtype *out) {
    return guard([&] {
        RC_CHECK(out, RC_ERR_INVALID_VALUE, "null out");
        dtype_size(ta); dtype_size(tb);
        *out = promote(ta, tb);
    });
}

int rc_binop_out_dtype_ex(rc_binop op, rc_dtype ta, rc_dtype tb, rc_dtype *out) {
    return guard([&] {
        RC_CHECK(out, RC_ERR_INVALID_VALUE, "null out");
        dtype_size(ta); dtype_size(tb);
        rc_dtype k;
        op_types(op, ta, tb, &k, out);
    });
}

int rc_op_mutc_refa_refb_ex(rc_device *dev, rc_binop op, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta,
                            const void *a, const rc_layout *la_, rc_dtype tb, const void *b, const rc_layout *lb_) {
    return guard([&] {
        DeviceGuard g(dev);
        rc_dtype k;
        check_out_type(op, tc, ta, tb, &k);
        if (op == RC_POW && pow_kind(ta, tb) != POW_SAME) {
            const rc_dtype te = pow_kind(ta, tb) == POW_FLOAT_INT ? RC_I32 : RC_U32;
            Layout lcc = from_c(lc), la = from_c(la_), lb = from_c(lb_);
            Operand ob;
            prepare(dev, te, tb, b, lb, &ob);
            CanonEw cn = canon_elementwise({&lcc, &la, &ob.l}, false, true);
            if (cn.empty) return;
            RC_CHECK(c && a && ob.ptr, RC_ERR_INVALID_VALUE, "null pointer");
            EwArgs args;
            args.c = c; args.a = a; args.b = ob.ptr;
            run_binary_pow_mixed(dev, ta, cn, args);
            return;
        }
        if (ta == k && tb == k) { status(rc_op_mutc_refa_refb(dev, op, k, c, lc, a, la_, b, lb_)); return; }
        {   // common pairs: the widening cast happens in registers (rc_ew_mixed.cu), one pass over memory
            Layout lcc = from_c(lc), la = from_c(la_), lb = from_c(lb_);
            CanonEw cn = canon_elementwise({&lcc, &la, &lb}, false, true);
            if (cn.empty) return;
            RC_CHECK(c && a && b, RC_ERR_INVALID_VALUE, "null pointer");
            EwArgs args;
            args.c = c; args.a = a; args.b = b;
            if (run_binary_promoted(dev, op, k, ta, tb, cn, args)) return;
        }
        Operand oa, ob;
        prepare(dev, k, ta, a, from_c(la_), &oa);
        prepare(dev, k, tb, b, from_c(lb_), &ob);
        rc_layout cla, clb;
        to_c(oa.l, &cla);
        to_c(ob.l, &clb);
        status(rc_op_mutc_refa_refb(dev, op, k, c, lc, oa.ptr, &cla, ob.ptr, &clb));
    });
}

int rc_op_mutc_refa_numb_ex(rc_device *dev, rc_binop op, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta,
                            const void *a, const rc_layout *la_, rc_dtype tb, const void *b_host) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(b_host != nullptr, RC_ERR_INVALID_VALUE, "null pointer: b");
        rc_dtype k;
        check_out_type(op, tc, ta, tb, &k);
        unsigned char sb[16];
        if (op == RC_POW && pow_kind(ta, tb) != POW_SAME) {
            cast_host_scalar(pow_kind(ta, tb) == POW_FLOAT_INT ? RC_I32 : RC_U32, tb, b_host, sb);
            Layout lcc = from_c(lc), la = from_c(la_);
            CanonEw cn = canon_elementwise({&lcc, &la}, false, true);
            if (cn.empty) return;
            RC_CHECK(c && a, RC_ERR_INVALID_VALUE, "null pointer");
            EwArgs args;
            args.c = c; args.a = a; args.b_const = true; args.b_host = sb;
            run_binary_pow_mixed(dev, ta, cn, args);
            return;
        }
        cast_host_scalar(k, tb, b_host, sb);
        if (ta != k) {
            Layout lcc = from_c(lc), la = from_c(la_);
            CanonEw cn = canon_elementwise({&lcc, &la}, false, true);
            if (cn.empty) return;
            RC_CHECK(c && a, RC_ERR_INVALID_VALUE, "null pointer");
            EwArgs args;
            args.c = c; args.a = a; args.b_const = true; args.b_host = sb;
            if (run_binary_promoted(dev, op, k, ta, k, cn, args)) return;
        }
        Operand oa;
        prepare(dev, k, ta, a, from_c(la_), &oa);
        rc_layout cla;
        to_c(oa.l, &cla);
        status(rc_op_mutc_refa_numb(dev, op, k, c, lc, oa.ptr, &cla, sb));
    });
}

int rc_op_mutc_numa_refb_ex(rc_device *dev, rc_binop op, rc_dtype tc, void *c, const rc_layout *lc, rc_dtype ta,
                            const void *a_host, rc_dtype tb, const void *b, const rc_layout *lb_) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(a_host != nullptr, RC_ERR_INVALID_VALUE, "null pointer: a");
        rc_dtype k;
        check_out_type(op, tc, ta, tb, &k);
        unsigned char sa[16];
        if (op == RC_POW && pow_kind(ta, tb) != POW_SAME) {
            const rc_dtype te = pow_kind(ta, tb) == POW_FLOAT_INT ? RC_I32 : RC_U32;
            Layout lcc = from_c(lc), lb = from_c(lb_);
            Operand ob;
            prepare(dev, te, tb, b, lb, &ob);
            CanonEw cn = canon_elementwise({&lcc, &ob.l}, false, true);
            if (cn.empty) return;
            RC_CHECK(c && ob.ptr, RC_ERR_INVALID_VALUE, "null pointer");
            EwArgs args;
            args.c = c; args.a_const = true; args.a_host = a_host; args.b = ob.ptr;
            run_binary_pow_mixed(dev, ta, cn, args);
            return;
        }
        cast_host_scalar(k, ta, a_host, sa);
        if (tb != k) {
            Layout lcc = from_c(lc), lb = from_c(lb_);
            CanonEw cn = canon_elementwise({&lcc, &lb}, false, true);
            if (cn.empty) return;
            RC_CHECK(c && b, RC_ERR_INVALID_VALUE, "null pointer");
            EwArgs args;
            args.c = c; args.a_const = true; args.a_host = sa; args.b = b;
            if (run_binary_promoted(dev, op, k, k, tb, cn, args)) return;
        }
        Operand ob;
        prepare(dev, k, tb, b, from_c(lb_), &ob);
        rc_layout clb;
        to_c(ob.l, &clb);
        status(rc_op_mutc_numa_refb(dev, op, k, c, lc, sa, ob.ptr, &clb));
    });
}

int rc_isclose(rc_device *dev, rc_dtype t, void *c, const rc_layout *lc_, const void *a, const rc_layout *la_,
               const void *b, const rc_layout *lb_, double rtol, double atol, int equal_nan) {
    return guard([&] {
        DeviceGuard g(dev);
        Layout lc = from_c(lc_), la = from_c(la_), lb = from_c(lb_);
        CanonEw cn = canon_elementwise({&lc, &la, &lb}, false, true);
        if (cn.empty) return;
        RC_CHECK(c && a && b, RC_ERR_INVALID_VALUE, "null pointer");
        IsCloseParams p;
        p.rtol = rtol; p.atol = atol; p.equal_nan = equal_nan ? 1 : 0;
        EwArgs args;
        args.c = c; args.a = a; args.b = b; args.params = &p;
        run_isclose(dev, t, cn, args);
    });
}

int rc_reduce_unraveled_arg_all(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const rc_layout *la_,
                                int64_t *index_out) {
    return guard([&] {
        RC_CHECK(op == RC_ARGMIN || op == RC_ARGMAX, RC_ERR_INVALID_VALUE, "unraveled arg takes RC_ARGMIN or RC_ARGMAX");
        Layout la = from_c(la_);
        RC_CHECK(index_out != nullptr || la.ndim() == 0, RC_ERR_INVALID_VALUE, "null index_out");
        uint64_t flat = 0;
        status(rc_reduce_all(dev, op, t, a, la_, &flat));  // row-major position over all axes; first occurrence wins
        for (int k = la.ndim() - 1; k >= 0; --k) {
            const uint64_t e = (uint64_t)la.shape[k];
            index_out[k] = (int64_t)(flat % e);
            flat /= e;
        }
    });
}

int rc_reduce_unraveled_arg_axes(rc_device *dev, rc_redop op, rc_dtype t, const void *a, const rc_layout *la_,
                                 const int64_t *axes_, int naxes, void **out_dev, rc_layout *lo_out) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(op == RC_ARGMIN || op == RC_ARGMAX, RC_ERR_INVALID_VALUE, "unraveled arg takes RC_ARGMIN or RC_ARGMAX");
        RC_CHECK(out_dev && lo_out, RC_ERR_INVALID_VALUE, "null out");
        *out_dev = nullptr;
        Layout la = from_c(la_);
        std::vector<int> axes = normalize_axes(axes_, naxes, la.ndim());
        void *flat = nullptr;
        status(rc_reduce_axes(dev, op, t, a, la_, axes_, naxes, &flat, lo_out));  // u64 positions, layout lo_out
        Layout lo = from_c(lo_out);
        int64_t mn = 0, mx = 0;
        bounds_index(lo, &mn, &mx);
        const int64_t count = std::max<int64_t>(mx, 1);  // rc_layout_for_reduce layouts are dense from offset 0
        void *out = nullptr;
        cudaError_t e = cudaMallocAsync(&out, (size_t)count * std::max(naxes, 1) * sizeof(uint64_t), dev->stream);
        if (e != cudaSuccess) {
            cudaFreeAsync(flat, dev->stream);
            raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
        }
        if (naxes > 0) {
            UnravelDesc d;
            d.n = naxes;
            for (int k = 0; k < naxes; ++k) d.shape[k] = la.shape[axes[k]];
            unravel_kernel<<<(unsigned)((count + 255) / 256), 256, 0, dev->stream>>>(
                d, static_cast<const uint64_t *>(flat), static_cast<uint64_t *>(out), count);
            after_launch(dev, "unravel_kernel");
        }
        cudaFreeAsync(flat, dev->stream);
        *out_dev = out;
    });
}

int rc_device_numa_node(const rc_device *dev, int *node) {
    return guard([&] {
        RC_CHECK(dev && node, RC_ERR_INVALID_VALUE, "null argument");
        *node = -1;
        char bus[32] = {0};
        if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev->ordinal) != cudaSuccess) { cudaGetLastError(); return; }
        std::string id(bus);
        std::transform(id.begin(), id.end(), id.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });
        std::ifstream f("/sys/bus/pci/devices/" + id + "/numa_node");
        int n = -1;
        if (f && (f >> n)) *node = n;
    });
}

int rc_host_alloc_on_node(size_t nbytes, int node, void **out, int *bound_out) {
    return guard([&] {
        RC_CHECK(out != nullptr, RC_ERR_INVALID_VALUE, "null out");
        if (bound_out) *bound_out = 0;
        if (node < 0) { status(rc_host_alloc(nbytes, out)); return; }
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        size_t bytes = ((nbytes ? nbytes : 1) + page - 1) / page * page;
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) raise(RC_ERR_MEMORY, "mmap failed for a pinned staging buffer");
        int bound = 0;
        if (node < 1024) {
            unsigned long mask[16] = {0};
            mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
            if (syscall(SYS_mbind, p, bytes, MPOL_BIND_, mask, (unsigned long)(8 * sizeof(mask) + 1), 0u) == 0) bound = 1;
        }
        if (!bound) {
            // mbind refused (containers without CAP_SYS_NICE filter it): first-touch the pages from a CPU of the node
            cpu_set_t saved, want;
            CPU_ZERO(&want);
            if (sched_getaffinity(0, sizeof(saved), &saved) == 0 && node_cpus(node, &want) &&
                sched_setaffinity(0, sizeof(want), &want) == 0) {
                volatile unsigned char *q = static_cast<volatile unsigned char *>(p);
                for (size_t off = 0; off < bytes; off += page) q[off] = 0;
                sched_setaffinity(0, sizeof(saved), &saved);
                bound = 2;
            }
        }
        // cudaHostRegister faults the pages in (on the bound node) and pins them
        cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
        if (e != cudaSuccess) {
            cudaGetLastError();
            munmap(p, bytes);
            raise(RC_ERR_MEMORY, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
        }
        {
            std::lock_guard<std::mutex> lock(g_numa_mu);
            g_numa_blocks[p] = bytes;
        }
        *out = p;
        if (bound_out) *bound_out = bound;
    });
}

}  // extern "C"
