// rc_copy.cu -- kernel family K2: strided copy / cast / fill.
//
// Replaces rstsr-native-impl/src/cpu_rayon/assignment.rs:14-225 (assign, assign_arbitary, fill and their
// *_promote variants).  Same-dtype copies move raw bits through unsigned integers of the element size, so
// they are bit-exact for every payload (NaN bit patterns included).  Copies whose two sides disagree on the
// fastest axis (to_contig of a permuted view) take ew_tile_kernel; all others ew_kernel.
#include "rc_dispatch.cuh"

namespace rc {

namespace {

template <class TO>
void cast_from(rc_device *dev, rc_dtype ta, const CanonEw &c, const EwArgs &args, bool out_bool) {
    // Casts between types of at most 4 bytes move < 9 bytes per element: on the scalar path they are element-rate bound
    // (~0.75 T elements/s: u8 -> f32 3.7 TB/s, i16 -> i32 4.4 TB/s), so those pairs get the pack kernels (32-byte packs on
    // the wide side).  Pairs with an 8-byte side are DRAM-bound on the scalar path already (6.5-7.1 TB/s) and stay there:
    // every vector instantiation costs three more kernels per pair.
#define RC_CAST_CASE(DT, CT, INB)                                                                   \
    case DT: {                                                                                      \
        constexpr bool VEC = sizeof(TO) <= 4 && sizeof(CT) <= 4;                                    \
        if (out_bool) ew_launch<FCast<TO, CT, true, INB>, false, false>(dev, c, args);              \
        else ew_launch<FCast<TO, CT, false, INB>, false, VEC>(dev, c, args);                        \
        return;                                                                                     \
    }
    switch (ta) {
        RC_CAST_CASE(RC_BOOL, uint8_t, true)
        RC_CAST_CASE(RC_I8, int8_t, false)
        RC_CAST_CASE(RC_I16, int16_t, false)
        RC_CAST_CASE(RC_I32, int32_t, false)
        RC_CAST_CASE(RC_I64, int64_t, false)
        RC_CAST_CASE(RC_U8, uint8_t, false)
        RC_CAST_CASE(RC_U16, uint16_t, false)
        RC_CAST_CASE(RC_U32, uint32_t, false)
        RC_CAST_CASE(RC_U64, uint64_t, false)
        RC_CAST_CASE(RC_F32, float, false)
        RC_CAST_CASE(RC_F64, double, false)
    }
#undef RC_CAST_CASE
    unsupported("cast from", ta);
}

// flattened-order copy with two independent index decompositions (no common refinement of the shapes)
struct ArbDesc {
    int ndim_c, ndim_a;
    uint32_t total;
    FastDiv div_c[RC_MAX_NDIM], div_a[RC_MAX_NDIM];
    int64_t stride_c[RC_MAX_NDIM], stride_a[RC_MAX_NDIM];
};

struct alignas(16) ArbWord16 { uint64_t lo, hi; };

template <class F>
__global__ void __launch_bounds__(256) arb_kernel(const __grid_constant__ ArbDesc d, typename F::TO *c, const typename F::TA *a) {
    uint32_t idx = blockIdx.x * 256u + threadIdx.x;
    if (idx >= d.total) return;
    int64_t oc = 0, oa = 0;
    uint32_t t = idx;
    for (int i = 0; i < d.ndim_c; ++i) {
        uint32_t q, r;
        d.div_c[i].divmod(t, q, r);
        oc += (int64_t)r * d.stride_c[i];
        t = q;
    }
    t = idx;
    for (int i = 0; i < d.ndim_a; ++i) {
        uint32_t q, r;
        d.div_a[i].divmod(t, q, r);
        oa += (int64_t)r * d.stride_a[i];
        t = q;
    }
    c[oc] = F::apply(a[oa]);
}

template <class F>
void arb_launch(rc_device *dev, const ArbDesc &d, void *c, const void *a) {
    uint32_t grid = (d.total + 255) / 256;
    arb_kernel<F><<<grid, 256, 0, dev->stream>>>(d, static_cast<typename F::TO *>(c),
                                                  static_cast<const typename F::TA *>(a));
    after_launch(dev, "arb_kernel");
}

}  // namespace

void run_cast(rc_device *dev, rc_dtype tc, rc_dtype ta, const CanonEw &c, const EwArgs &args) {
    if (tc == ta) {
        switch (dtype_size(tc)) {
            case 1: ew_launch<FIdentity<uint8_t>>(dev, c, args); return;
            case 2: ew_launch<FIdentity<uint16_t>>(dev, c, args); return;
            case 4: ew_launch<FIdentity<uint32_t>>(dev, c, args); return;
            case 8: ew_launch<FIdentity<uint64_t>>(dev, c, args); return;
            case 16: run_copy16(dev, c, args); return;
        }
    }
    if (dtype_is_extended(tc) || dtype_is_extended(ta)) {
        if (run_cast_ext(dev, tc, ta, c, args)) return;
        // no direct kernel: stage through the real type the conversion is defined by.  primitive -> Complex<R> is
        // `Complex::new(self as R, 0)` (DTypeCastAPI, promotion.rs:453-458); integer <-> half goes through f64 / f32 as
        // half::f16::from_f64(v as f64) / (h.to_f32() as int) would.  One extra pass over a compact temporary.
        rc_dtype via = RC_BOOL;
        if (dtype_is_complex(tc) && !dtype_is_extended(ta)) via = tc == RC_C32 ? RC_F32 : RC_F64;
        else if (dtype_is_half(tc) && !dtype_is_extended(ta)) via = RC_F64;
        else if (dtype_is_half(ta) && !dtype_is_extended(tc)) via = RC_F32;
        if (via != RC_BOOL) {
            CanonEw c1 = c, c2 = c;  // slot 0 = output, slot 1 = source; the temporary is contiguous over the canonical shape
            int64_t n = 1;
            for (int i = 0; i < c.ndim; ++i) {
                c1.stride[0][i] = n;
                c2.stride[1][i] = n;
                n *= c.shape[i];
            }
            c1.base[0] = 0;
            c2.base[1] = 0;
            void *tmp = nullptr;
            cudaError_t e = cudaMallocAsync(&tmp, (size_t)n * dtype_size(via), dev->stream);
            if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
            try {
                EwArgs a1 = args, a2 = args;
                a1.c = tmp;
                a2.a = tmp;
                run_cast(dev, via, ta, c1, a1);
                run_cast(dev, tc, via, c2, a2);
            } catch (...) {
                cudaFreeAsync(tmp, dev->stream);
                throw;
            }
            cudaFreeAsync(tmp, dev->stream);
            return;
        }
        raise(RC_ERR_UNIMPLEMENTED, std::string("cast ") + dtype_name(ta) + " -> " + dtype_name(tc) + " is not implemented");
    }
    switch (tc) {
        case RC_BOOL: cast_from<uint8_t>(dev, ta, c, args, true); return;
        case RC_I8: cast_from<int8_t>(dev, ta, c, args, false); return;
        case RC_I16: cast_from<int16_t>(dev, ta, c, args, false); return;
        case RC_I32: cast_from<int32_t>(dev, ta, c, args, false); return;
        case RC_I64: cast_from<int64_t>(dev, ta, c, args, false); return;
        case RC_U8: cast_from<uint8_t>(dev, ta, c, args, false); return;
        case RC_U16: cast_from<uint16_t>(dev, ta, c, args, false); return;
        case RC_U32: cast_from<uint32_t>(dev, ta, c, args, false); return;
        case RC_U64: cast_from<uint64_t>(dev, ta, c, args, false); return;
        case RC_F32: cast_from<float>(dev, ta, c, args, false); return;
        case RC_F64: cast_from<double>(dev, ta, c, args, false); return;
    }
    unsupported("cast to", tc);
}

void run_fill(rc_device *dev, rc_dtype tc, const CanonEw &c, void *c_ptr, const void *value_tc) {
    EwArgs args;
    args.c = c_ptr;
    args.a_host = value_tc;
    switch (dtype_size(tc)) {
        case 1: ew_launch<FFill<uint8_t>, false>(dev, c, args); return;
        case 2: ew_launch<FFill<uint16_t>, false>(dev, c, args); return;
        case 4: ew_launch<FFill<uint32_t>, false>(dev, c, args); return;
        case 8: ew_launch<FFill<uint64_t>, false>(dev, c, args); return;
        case 16: run_fill16(dev, c, args); return;
    }
    unsupported("fill", tc);
}

void run_assign_arbitrary_generic(rc_device *dev, rc_dtype tc, void *c, const Layout &lc, rc_dtype ta, const void *a,
                                  const Layout &la, rc_order order) {
    int64_t total = lc.size();
    if (total == 0) return;
    RC_CHECK(total <= kMaxItemsPerLaunch, RC_ERR_UNIMPLEMENTED,
             "assign_arbitary between shapes without a common refinement is limited to 2^31 elements");
    ArbDesc d;
    std::memset(&d, 0, sizeof(d));
    d.total = (uint32_t)total;
    auto fill = [&](const Layout &l, int *nd, FastDiv *dv, int64_t *st) {
        int n = l.ndim();
        *nd = n;
        for (int k = 0; k < n; ++k) {
            int i = (order == RC_ROW_MAJOR) ? n - 1 - k : k;  // fastest axis first
            dv[k] = FastDiv((uint32_t)l.shape[i]);
            st[k] = l.stride[i];
        }
    };
    fill(lc, &d.ndim_c, d.div_c, d.stride_c);
    fill(la, &d.ndim_a, d.div_a, d.stride_a);
    RC_CHECK(tc == ta, RC_ERR_RUNTIME, "generic flattened-order copy takes one dtype (casts go through a staging copy)");
    void *pc = static_cast<char *>(c) + lc.offset * (int64_t)dtype_size(tc);
    const void *pa = static_cast<const char *>(a) + la.offset * (int64_t)dtype_size(ta);
    switch (dtype_size(tc)) {
        case 1: arb_launch<FIdentity<uint8_t>>(dev, d, pc, pa); return;
        case 2: arb_launch<FIdentity<uint16_t>>(dev, d, pc, pa); return;
        case 4: arb_launch<FIdentity<uint32_t>>(dev, d, pc, pa); return;
        case 8: arb_launch<FIdentity<uint64_t>>(dev, d, pc, pa); return;
        case 16: arb_launch<FIdentity<ArbWord16>>(dev, d, pc, pa); return;
    }
    unsupported("copy", tc);
}

}  // namespace rc
