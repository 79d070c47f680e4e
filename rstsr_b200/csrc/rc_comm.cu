// rc_comm.cu -- the one exchange step of the path: combining per-GPU partials of a sharded reduction.
//
// Two transports behind one entry point:
//   * PEER WINDOW (results up to 256 KiB): every rank owns a small cudaMalloc'd window that all other ranks map
//     (CUDA IPC between processes, plain peer access inside one process).  ONE kernel per rank (a) folds the local
//     partial states of the two-pass reduction -- i.e. it IS the reduction's second pass --, (b) stores the folded
//     values into every rank's window over NVLink, (c) publishes a per-block epoch flag with st.release.sys,
//     (d) waits for the same block of every other rank with ld.acquire.sys and (e) folds the N contributions in RANK
//     ORDER.  The result is bitwise identical on every rank and run-to-run (NCCL's choice of ring / tree / NVLS is
//     not), and costs one launch instead of "second pass + ncclAllReduce" (~25-35 us for 128 KiB at N = 8).
//   * NCCL all-reduce for anything larger (bandwidth-bound; NVLS / ring over NVSwitch is the right tool there).
// NCCL is bound at run time (dlopen of libnccl.so.2, which resolves to the copy torch already loaded when the caller
// is a torch process), so the library has no link-time dependency and single-GPU users never touch it.  NCCL also
// bootstraps the window exchange (all-gather of the IPC handles).
//
// Combiners follow the reference's closures (rstsr-core/src/feature_rayon/auto_impl/reduction.rs:14-63, 94-107,
// 138-185): sum 0,+ ; prod 1,* ; max T::MIN, ext_max ; min T::MAX, ext_min ; mean = sum, then / T::from_usize(n)
// with n the GLOBAL reduced count.  An empty local shard contributes the monoid identity (every rank always takes
// part in the exchange -- a rank-local error before a collective would hang the others).
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include "rc_layout.hpp"
#include "rc_reduce.cuh"

namespace rc {
namespace {

constexpr int PEER_MAX_RANKS = 16;
constexpr int PEER_BLOCK = 256;
constexpr int PEER_MAX_BLOCKS = 128;
constexpr size_t PEER_SLOT_BYTES = 256u << 10;  // per source rank and parity
constexpr size_t PEER_FLAG_BYTES = 2ull * PEER_MAX_RANKS * PEER_MAX_BLOCKS * sizeof(unsigned long long);

}  // namespace
}  // namespace rc

struct rc_comm {
    rc_device *dev = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // peer window
    bool peer = false;
    unsigned char *win_local = nullptr;
    unsigned char *win[rc::PEER_MAX_RANKS] = {};
    bool ipc_opened[rc::PEER_MAX_RANKS] = {};
    unsigned long long epoch = 0;     // collective calls made through the window (same on every rank)
    void *scratch = nullptr;          // 64 device bytes: local scalar of a sharded *_all when it took one pass
    std::mutex mu;
};

namespace rc {
namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    });
    RC_CHECK(api.handle && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather,
             RC_ERR_DEVICE, "NCCL (libnccl.so.2) could not be loaded");
    return api;
}

void nccl_check(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return;
    NcclApi &api = nccl();
    raise(RC_ERR_DEVICE, std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(r) : "NCCL error"));
}

// NCCL has no 16-bit integers: i16 / u16 travel through the peer window, or widened (see all_reduce_nccl)
bool nccl_has_dtype(rc_dtype t) { return t != RC_I16 && t != RC_U16; }

ncclDataType_t nccl_dtype(rc_dtype t) {
    switch (t) {
        case RC_I8: return ncclInt8;
        case RC_U8: case RC_BOOL: return ncclUint8;
        case RC_I32: return ncclInt32;
        case RC_U32: return ncclUint32;
        case RC_I64: return ncclInt64;
        case RC_U64: return ncclUint64;
        case RC_F32: return ncclFloat32;
        case RC_F64: return ncclFloat64;
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, std::string("all-reduce is not implemented for dtype ") + dtype_name(t));
}

ncclRedOp_t nccl_op(rc_redop op) {
    switch (op) {
        case RC_SUM: case RC_MEAN: return ncclSum;
        case RC_PROD: return ncclProd;
        case RC_MAX: return ncclMax;
        case RC_MIN: return ncclMin;
        default: break;
    }
    raise(RC_ERR_INVALID_VALUE, "unknown reduction op");
}

// ------------------------------------------------------------------------------------------------
// peer window kernel
// ------------------------------------------------------------------------------------------------
struct PeerArgs {
    unsigned char *win[PEER_MAX_RANKS];
    int nranks, rank;
    unsigned long long epoch;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Window layout: flags u64[2 parity][PEER_MAX_RANKS source][PEER_MAX_BLOCKS], then data [2 parity][nranks source]
// [PEER_SLOT_BYTES].  Double buffering by epoch parity is enough: a rank can start call e+2 (same parity as e) only
// after every rank has published call e+1, which each does after it finished reading call e.
//
// opb = outputs per block (power of two <= 256), a function of `count` ONLY, so block b owns the same outputs on
// every rank whatever its local split factor S is; the 256 / opb threads of a group fold the S local states of one
// output in a fixed tree.
template <class P>
__global__ void __launch_bounds__(PEER_BLOCK) peer_combine_kernel(const __grid_constant__ PeerArgs pa,
                                                                  const typename P::S *__restrict__ partial, int64_t S,
                                                                  int64_t pitch, int64_t count, typename P::S *out,
                                                                  int opb, int mean, int64_t n_div,
                                                                  unsigned long long timeout_ns) {
    using T = typename P::S;
    __shared__ T warp_acc[PEER_BLOCK / 32];
    const int tid = threadIdx.x;
    const int tpo = PEER_BLOCK / opb;
    const int g = tid / tpo, t = tid - g * tpo;
    const int par = (int)(pa.epoch & 1ull);
    const size_t my_slot = PEER_FLAG_BYTES + ((size_t)par * pa.nranks + pa.rank) * PEER_SLOT_BYTES;

    // (a) + (b): fold the local states, store the value into every rank's window
    for (int64_t base = (int64_t)blockIdx.x * opb; base < count; base += (int64_t)gridDim.x * opb) {
        const int64_t o = base + g;
        T v = P::init();
        if (o < count)
            for (int64_t s = t; s < S; s += tpo) v = P::comb(v, partial[s * pitch + o]);
        if (tpo <= 32) {
            for (int m = tpo >> 1; m >= 1; m >>= 1) v = P::comb(v, shfl_xor_state<T>(v, m));
        } else {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) v = P::comb(v, shfl_xor_state<T>(v, m));
            __syncthreads();  // warp_acc may still be read by the previous iteration
            if ((tid & 31) == 0) warp_acc[tid >> 5] = v;
            __syncthreads();
            if (t == 0) {
                const int w0 = tid >> 5, nw = tpo >> 5;
                v = warp_acc[w0];
                for (int w = 1; w < nw; ++w) v = P::comb(v, warp_acc[w0 + w]);
            }
        }
        if (o < count && t == 0) {
            for (int r = 0; r < pa.nranks; ++r) reinterpret_cast<T *>(pa.win[r] + my_slot)[o] = v;
        }
    }
    __syncthreads();
    // (c) publish: one thread per destination rank; the fence orders the block's stores (observed through the
    // barrier) before the flag at system scope
    if (tid < pa.nranks) {
        __threadfence_system();
        unsigned long long *f = reinterpret_cast<unsigned long long *>(pa.win[tid]) +
                                ((size_t)par * PEER_MAX_RANKS + pa.rank) * PEER_MAX_BLOCKS + blockIdx.x;
        st_release_sys(f, pa.epoch);
    }
    // (d) wait for block blockIdx.x of every rank
    if (tid < pa.nranks) {
        const unsigned long long *f = reinterpret_cast<const unsigned long long *>(pa.win[pa.rank]) +
                                      ((size_t)par * PEER_MAX_RANKS + tid) * PEER_MAX_BLOCKS + blockIdx.x;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) != pa.epoch) {
            if (global_timer_ns() - t0 > timeout_ns) {
                printf("rstsr-cuda: peer combine timed out (rank %d waits for rank %d, block %d, epoch %llu)\n", pa.rank,
                       tid, (int)blockIdx.x, pa.epoch);
                __trap();  // a lost rank must not hang the job
            }
        }
    }
    __syncthreads();
    // (e) fold the N contributions in rank order (volatile: the lines were written by remote GPUs)
    const unsigned char *mine = pa.win[pa.rank] + PEER_FLAG_BYTES + (size_t)par * pa.nranks * PEER_SLOT_BYTES;
    for (int64_t base = (int64_t)blockIdx.x * opb; base < count; base += (int64_t)gridDim.x * opb) {
        const int64_t o = base + tid;
        if (tid < opb && o < count) {
            T v = reinterpret_cast<const volatile T *>(mine)[o];
            for (int r = 1; r < pa.nranks; ++r)
                v = P::comb(v, (T) reinterpret_cast<const volatile T *>(mine + (size_t)r * PEER_SLOT_BYTES)[o]);
            if constexpr (std::is_floating_point<T>::value) {
                if (mean) v = v / (T)n_div;
            }
            out[o] = v;
        }
    }
}

template <class P>
__global__ void identity_kernel(typename P::S *out, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = P::init();
}

unsigned long long peer_timeout_ns() {
    static unsigned long long v = [] {
        const char *e = getenv("RC_COMM_TIMEOUT_S");
        double s = e ? atof(e) : 120.0;
        if (!(s > 0)) s = 120.0;
        return (unsigned long long)(s * 1e9);
    }();
    return v;
}

int peer_opb(int64_t count, int sm_count) {
    // outputs per block: enough blocks to spread the stores (<= PEER_MAX_BLOCKS, all co-resident), few enough
    // outputs per block that small results get many threads per output
    int64_t blocks = std::min<int64_t>(PEER_MAX_BLOCKS, std::max(1, sm_count - 20));
    int64_t want = (count + blocks - 1) / blocks;
    int opb = 1;
    while (opb < want && opb < PEER_BLOCK) opb *= 2;
    return opb;
}

template <class P>
void launch_peer(rc_comm *c, const void *partial, int64_t S, int64_t pitch, int64_t count, void *out, bool mean,
                 int64_t n_div) {
    using T = typename P::S;
    PeerArgs pa;
    for (int r = 0; r < PEER_MAX_RANKS; ++r) pa.win[r] = c->win[r];
    pa.nranks = c->nranks;
    pa.rank = c->rank;
    pa.epoch = ++c->epoch;
    // identical on every rank: depends on `count` and a constant only (ranks of one job share the GPU model)
    const int opb = peer_opb(count, 148);
    int64_t blocks = std::min<int64_t>(PEER_MAX_BLOCKS, (count + opb - 1) / opb);
    if (blocks < 1) blocks = 1;
    peer_combine_kernel<P><<<(unsigned)blocks, PEER_BLOCK, 0, c->dev->stream>>>(
        pa, static_cast<const T *>(partial), S, pitch, count, static_cast<T *>(out), opb, mean ? 1 : 0, n_div,
        peer_timeout_ns());
    after_launch(c->dev, "peer_combine_kernel");
}

template <class P>
void launch_identity(rc_device *dev, void *out, int64_t count) {
    if (count <= 0) return;
    identity_kernel<P><<<(unsigned)((count + 255) / 256), 256, 0, dev->stream>>>(static_cast<typename P::S *>(out), count);
    after_launch(dev, "identity_kernel");
}

// op x dtype dispatch of a generic lambda taking a policy tag
template <class P> struct Tag { using type = P; };

template <class T, class F>
void with_policy_t(rc_redop op, F &&f) {
    switch (op) {
        case RC_SUM: case RC_MEAN: f(Tag<PSum<T>>()); return;
        case RC_PROD: f(Tag<PProd<T>>()); return;
        case RC_MAX: f(Tag<PMax<T>>()); return;
        case RC_MIN: f(Tag<PMin<T>>()); return;
        default: break;
    }
    raise(RC_ERR_UNIMPLEMENTED, "sharded reductions cover sum / prod / max / min / mean");
}

template <class F>
void with_policy(rc_redop op, rc_dtype t, F &&f) {
    switch (t) {
        case RC_I8: with_policy_t<int8_t>(op, f); return;
        case RC_U8: case RC_BOOL: with_policy_t<uint8_t>(op, f); return;
        case RC_I16: with_policy_t<int16_t>(op, f); return;
        case RC_U16: with_policy_t<uint16_t>(op, f); return;
        case RC_I32: with_policy_t<int32_t>(op, f); return;
        case RC_U32: with_policy_t<uint32_t>(op, f); return;
        case RC_I64: with_policy_t<int64_t>(op, f); return;
        case RC_U64: with_policy_t<uint64_t>(op, f); return;
        case RC_F32: with_policy_t<float>(op, f); return;
        case RC_F64: with_policy_t<double>(op, f); return;
    }
    raise(RC_ERR_INVALID_VALUE, "unknown dtype");
}

bool peer_fits(const rc_comm *c, rc_dtype t, int64_t count) {
    return c->peer && (size_t)count * dtype_size(t) <= PEER_SLOT_BYTES;
}

// out[i] = combine over ranks of (fold over s < S of partial[s * pitch + i]); mean divides by n_div afterwards
void combine_peer(rc_comm *c, rc_redop op, rc_dtype t, const void *partial, int64_t S, int64_t pitch, int64_t count,
                  void *out, int64_t n_div) {
    with_policy(op, t, [&](auto tag) {
        using P = typename decltype(tag)::type;
        launch_peer<P>(c, partial, S, pitch, count, out, op == RC_MEAN, n_div);
    });
}

void fill_identity(rc_device *dev, rc_redop op, rc_dtype t, void *out, int64_t count) {
    with_policy(op, t, [&](auto tag) {
        using P = typename decltype(tag)::type;
        launch_identity<P>(dev, out, count);
    });
}

// in-place NCCL all-reduce of a dense device block; 16-bit integers are widened to 32 bits (wrapping sums and
// products commute with truncation, max / min are order-preserving)
void all_reduce_nccl(rc_comm *c, rc_redop op, rc_dtype t, void *buf, size_t count) {
    rc_device *dev = c->dev;
    if (nccl_has_dtype(t)) {
        nccl_check(nccl().AllReduce(buf, buf, count, nccl_dtype(t), nccl_op(op), c->comm, dev->stream), "ncclAllReduce");
        return;
    }
    const rc_dtype wide = (t == RC_I16) ? RC_I32 : RC_U32;
    void *tmp = nullptr;
    cudaError_t e = cudaMallocAsync(&tmp, count * 4, dev->stream);
    if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
    rc_layout l;
    std::memset(&l, 0, sizeof(l));
    l.ndim = 1; l.shape[0] = (int64_t)count; l.stride[0] = 1;
    int st = rc_assign(dev, wide, tmp, &l, t, buf, &l);
    ncclResult_t nr = ncclSuccess;
    if (st == RC_OK) nr = nccl().AllReduce(tmp, tmp, count, nccl_dtype(wide), nccl_op(op), c->comm, dev->stream);
    if (st == RC_OK && nr == ncclSuccess) st = rc_assign(dev, t, buf, &l, wide, tmp, &l);
    cudaFreeAsync(tmp, dev->stream);
    nccl_check(nr, "ncclAllReduce");
    if (st != RC_OK) raise((rc_status)st, rc_last_error());
}

// mean after an NCCL sum: divide in place by the global count (T::from_usize(n))
void divide_in_place(rc_device *dev, rc_dtype t, void *buf, int64_t count, int64_t n_div) {
    RC_CHECK(dtype_is_float(t), RC_ERR_UNIMPLEMENTED, "mean requires a floating-point dtype");
    rc_layout l;
    std::memset(&l, 0, sizeof(l));
    l.ndim = 1; l.shape[0] = count; l.stride[0] = 1;
    double d64 = (double)n_div;
    float d32 = (float)n_div;
    int st = rc_op_muta_numb(dev, RC_DIV, t, buf, &l, t == RC_F64 ? (const void *)&d64 : (const void *)&d32, 0);
    if (st != RC_OK) raise((rc_status)st, rc_last_error());
}

// ------------------------------------------------------------------------------------------------
// window set-up: every rank allocates, all-gathers {ipc handle, pid, pointer, ordinal}, maps the others
// ------------------------------------------------------------------------------------------------
struct PeerInfo {
    cudaIpcMemHandle_t handle;
    uint64_t pid;
    uint64_t ptr;
    int32_t ordinal;
    int32_t ok;
    char host[64];
};

void peer_teardown(rc_comm *c) {
    for (int r = 0; r < c->nranks && r < PEER_MAX_RANKS; ++r) {
        if (c->ipc_opened[r] && c->win[r]) cudaIpcCloseMemHandle(c->win[r]);
        c->win[r] = nullptr;
        c->ipc_opened[r] = false;
    }
    if (c->win_local) cudaFree(c->win_local);
    c->win_local = nullptr;
    c->peer = false;
}

void peer_setup(rc_comm *c) {
    rc_device *dev = c->dev;
    const char *env = getenv("RC_COMM_PEER");
    int ok = (env && atoi(env) == 0) ? 0 : 1;
    if (c->nranks > PEER_MAX_RANKS) ok = 0;
    const size_t bytes = PEER_FLAG_BYTES + 2ull * c->nranks * PEER_SLOT_BYTES;
    PeerInfo mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok) {
        if (cudaMalloc((void **)&c->win_local, bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; c->win_local = nullptr; }
    }
    if (ok) {
        if (cudaMemsetAsync(c->win_local, 0, bytes, dev->stream) != cudaSuccess ||
            cudaStreamSynchronize(dev->stream) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    if (ok && c->nranks > 1) {
        if (cudaIpcGetMemHandle(&mine.handle, c->win_local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    mine.pid = (uint64_t)getpid();
    mine.ptr = (uint64_t)(uintptr_t)c->win_local;
    mine.ordinal = dev->ordinal;
    mine.ok = ok;
    gethostname(mine.host, sizeof(mine.host) - 1);

    // all-gather the infos through a device buffer (every rank takes part whatever its local `ok` is)
    std::vector<PeerInfo> all((size_t)c->nranks);
    PeerInfo *dbuf = nullptr;
    RC_CUDA(cudaMalloc((void **)&dbuf, sizeof(PeerInfo) * c->nranks));
    try {
        RC_CUDA(cudaMemcpyAsync(dbuf + c->rank, &mine, sizeof(PeerInfo), cudaMemcpyHostToDevice, dev->stream));
        nccl_check(nccl().AllGather(dbuf + c->rank, dbuf, sizeof(PeerInfo), ncclUint8, c->comm, dev->stream), "ncclAllGather");
        RC_CUDA(cudaMemcpyAsync(all.data(), dbuf, sizeof(PeerInfo) * c->nranks, cudaMemcpyDeviceToHost, dev->stream));
        RC_CUDA(cudaStreamSynchronize(dev->stream));
        for (int r = 0; r < c->nranks; ++r) ok = ok && all[r].ok;
        for (int r = 0; r < c->nranks && ok; ++r) {
            if (r == c->rank) { c->win[r] = c->win_local; continue; }
            if (std::strncmp(all[r].host, mine.host, sizeof(mine.host)) != 0) { ok = 0; break; }  // one node only
            if (all[r].pid == mine.pid) {
                // another handle of this process: plain peer access
                if (all[r].ordinal != dev->ordinal) {
                    int can = 0;
                    if (cudaDeviceCanAccessPeer(&can, dev->ordinal, all[r].ordinal) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
                    cudaError_t e = cudaDeviceEnablePeerAccess(all[r].ordinal, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = 0; break; }
                    cudaGetLastError();
                }
                c->win[r] = (unsigned char *)(uintptr_t)all[r].ptr;
            } else {
                void *p = nullptr;
                if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
                c->win[r] = (unsigned char *)p;
                c->ipc_opened[r] = true;
            }
        }
        // agree: the window is used only if EVERY rank mapped every other rank (a mixed choice would hang)
        int32_t flag = ok;
        RC_CUDA(cudaMemcpyAsync(dbuf, &flag, sizeof(flag), cudaMemcpyHostToDevice, dev->stream));
        nccl_check(nccl().AllReduce(dbuf, dbuf, 1, ncclInt32, ncclMin, c->comm, dev->stream), "ncclAllReduce");
        RC_CUDA(cudaMemcpyAsync(&flag, dbuf, sizeof(flag), cudaMemcpyDeviceToHost, dev->stream));
        RC_CUDA(cudaStreamSynchronize(dev->stream));
        ok = flag;
    } catch (...) {
        cudaFree(dbuf);
        peer_teardown(c);
        throw;
    }
    cudaFree(dbuf);
    if (!ok) { peer_teardown(c); return; }
    c->peer = true;
}

// dense block [lo, hi) of element offsets covered by a layout whose elements are all distinct and gap-free
bool dense_block(const Layout &l, int64_t *first) {
    int64_t mn = 0, mx = 0;
    bounds_index(l, &mn, &mx);
    for (auto s : l.stride) if (s == 0) return false;
    if (mx - mn != l.size()) return false;
    *first = mn;
    return true;
}

}  // namespace
}  // namespace rc

using namespace rc;

extern "C" {

int rc_comm_get_unique_id(uint8_t id[RC_COMM_ID_BYTES]) {
    return guard([&] {
        static_assert(sizeof(ncclUniqueId) == RC_COMM_ID_BYTES, "ncclUniqueId size");
        RC_CHECK(id != nullptr, RC_ERR_INVALID_VALUE, "null id");
        ncclUniqueId u;
        nccl_check(nccl().GetUniqueId(&u), "ncclGetUniqueId");
        std::memcpy(id, &u, sizeof(u));
    });
}

int rc_comm_init_rank(rc_device *dev, int nranks, int rank, const uint8_t id[RC_COMM_ID_BYTES], rc_comm **out) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(out && id, RC_ERR_INVALID_VALUE, "null argument");
        RC_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, RC_ERR_INVALID_VALUE, "invalid rank / nranks");
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof(u));
        std::unique_ptr<rc_comm> c(new rc_comm());
        c->dev = dev;
        c->nranks = nranks;
        c->rank = rank;
        nccl_check(nccl().CommInitRank(&c->comm, nranks, u, rank), "ncclCommInitRank");
        try {
            RC_CUDA(cudaMalloc(&c->scratch, 64));
            peer_setup(c.get());
        } catch (...) {
            if (c->scratch) cudaFree(c->scratch);
            nccl().CommDestroy(c->comm);
            throw;
        }
        *out = c.release();
    });
}

int rc_comm_destroy(rc_comm *comm) {
    return guard([&] {
        if (!comm) return;
        cudaSetDevice(comm->dev->ordinal);
        cudaStreamSynchronize(comm->dev->stream);
        peer_teardown(comm);
        if (comm->scratch) cudaFree(comm->scratch);
        if (comm->comm) nccl().CommDestroy(comm->comm);
        delete comm;
    });
}

int rc_comm_info(const rc_comm *comm, int *nranks, int *rank, int *peer_window) {
    return guard([&] {
        RC_CHECK(comm != nullptr, RC_ERR_INVALID_VALUE, "null comm");
        if (nranks) *nranks = comm->nranks;
        if (rank) *rank = comm->rank;
        if (peer_window) *peer_window = comm->peer ? 1 : 0;
    });
}

int rc_comm_set_peer_window(rc_comm *comm, int enable) {
    return guard([&] {
        RC_CHECK(comm != nullptr, RC_ERR_INVALID_VALUE, "null comm");
        std::lock_guard<std::mutex> lock(comm->mu);
        if (enable) RC_CHECK(comm->win_local != nullptr && comm->win[comm->rank] != nullptr, RC_ERR_DEVICE,
                             "the peer window was not mapped when the communicator was created");
        comm->peer = enable != 0;
    });
}

int rc_comm_all_reduce(rc_comm *comm, rc_redop op, rc_dtype t, void *buf, size_t count) {
    return guard([&] {
        RC_CHECK(comm != nullptr, RC_ERR_INVALID_VALUE, "null comm");
        DeviceGuard g(comm->dev);
        RC_CHECK(op <= RC_MEAN, RC_ERR_UNIMPLEMENTED, "all-reduce covers sum / prod / max / min (mean = sum)");
        if (count == 0) return;
        RC_CHECK(buf != nullptr, RC_ERR_INVALID_VALUE, "null buffer");
        std::lock_guard<std::mutex> lock(comm->mu);
        const rc_redop comb = (op == RC_MEAN) ? RC_SUM : op;
        if (peer_fits(comm, t, (int64_t)count)) combine_peer(comm, comb, t, buf, 1, 0, (int64_t)count, buf, 1);
        else all_reduce_nccl(comm, comb, t, buf, count);
    });
}

int rc_reduce_all_sharded(rc_device *dev, rc_comm *comm, rc_redop op, rc_dtype t, const void *a, const rc_layout *la_,
                          int64_t n_global, void *host_out) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(comm != nullptr && comm->dev == dev, RC_ERR_DEVICE_MISMATCH, "communicator belongs to another device");
        RC_CHECK(host_out != nullptr, RC_ERR_INVALID_VALUE, "null host_out");
        RC_CHECK(op <= RC_MEAN, RC_ERR_UNIMPLEMENTED, "sharded reductions cover sum / prod / max / min / mean");
        if (op == RC_MEAN) RC_CHECK(dtype_is_float(t), RC_ERR_UNIMPLEMENTED, "mean requires a floating-point dtype");
        // every check below is a function of the arguments all ranks share, so no rank can leave the others waiting
        if (op == RC_MAX || op == RC_MIN)
            RC_CHECK(n_global != 0, RC_ERR_INVALID_VALUE,
                     op == RC_MAX ? "zero-size array is not supported for max" : "zero-size array is not supported for min");
        Layout la = from_c(la_);
        if (la.size() != 0) RC_CHECK(a != nullptr, RC_ERR_INVALID_VALUE, "null pointer: a");
        const rc_redop local = (op == RC_MEAN) ? RC_SUM : op;  // mean: sum, divided by the GLOBAL count at the end
        std::vector<int> axes(la.ndim());
        for (int i = 0; i < la.ndim(); ++i) axes[i] = i;
        Layout lo;  // 0-d

        std::lock_guard<std::mutex> clock(comm->mu);
        std::lock_guard<std::mutex> slot_lock(dev->slot_mu);
        void *host = nullptr;
        void *slot = scalar_slot(dev, &host);
        {
            std::lock_guard<std::mutex> ws_lock(dev->ws_mu);
            const bool peer = peer_fits(comm, t, 1);
            dev->preq = rc_device::PartialReq();
            dev->preq.want = peer;
            const bool empty_local = la.size() == 0;
            if (!empty_local) reduce_local(dev, local, t, a, la, axes, comm->scratch, lo);
            else fill_identity(dev, local, t, comm->scratch, 1);
            const rc_device::PartialReq rq = dev->preq;
            dev->preq = rc_device::PartialReq();
            if (peer) {
                // the combine kernel writes the final scalar straight into the handle's mapped host slot
                if (rq.got) combine_peer(comm, op, t, rq.ptr, rq.S, rq.pitch, 1, slot, n_global);
                else combine_peer(comm, op, t, comm->scratch, 1, 0, 1, slot, n_global);
            } else {
                all_reduce_nccl(comm, local, t, comm->scratch, 1);
                if (op == RC_MEAN) divide_in_place(dev, t, comm->scratch, 1, n_global);
                RC_CUDA(cudaMemcpyAsync(host, comm->scratch, dtype_size(t), cudaMemcpyDeviceToHost, dev->stream));
            }
        }
        RC_CUDA(cudaStreamSynchronize(dev->stream));
        std::memcpy(host_out, host, dtype_size(t));
    });
}

int rc_reduce_axes_sharded(rc_device *dev, rc_comm *comm, rc_redop op, rc_dtype t, const void *a, const rc_layout *la_,
                           const int64_t *axes_, int naxes, int64_t n_reduced_global, void *out, const rc_layout *lo_) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(comm != nullptr && comm->dev == dev, RC_ERR_DEVICE_MISMATCH, "communicator belongs to another device");
        RC_CHECK(op <= RC_MEAN, RC_ERR_UNIMPLEMENTED, "sharded reductions cover sum / prod / max / min / mean");
        if (op == RC_MEAN) RC_CHECK(dtype_is_float(t), RC_ERR_UNIMPLEMENTED, "mean requires a floating-point dtype");
        if (op == RC_MAX || op == RC_MIN)
            RC_CHECK(n_reduced_global != 0, RC_ERR_INVALID_VALUE,
                     op == RC_MAX ? "zero-size array is not supported for max" : "zero-size array is not supported for min");
        Layout la = from_c(la_), lo = from_c(lo_);
        std::vector<int> axes = normalize_axes(axes_, naxes, la.ndim());
        const int64_t count = lo.size();
        if (count == 0) return;
        RC_CHECK(out != nullptr, RC_ERR_INVALID_VALUE, "null pointer: out");
        const rc_redop local = (op == RC_MEAN) ? RC_SUM : op;
        int64_t first = 0;
        const bool dense = dense_block(lo, &first);
        const size_t es = dtype_size(t);

        std::lock_guard<std::mutex> clock(comm->mu);
        // a strided / broadcast output is reduced into a dense temporary, combined there and assigned back
        void *tmp = nullptr;
        Layout ltmp = lo;
        void *dst = out;
        if (!dense) {
            ltmp = new_contig(lo.shape, RC_ROW_MAJOR, 0);
            cudaError_t e = cudaMallocAsync(&tmp, (size_t)count * es, dev->stream);
            if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
            dst = tmp;
            first = 0;
        }
        try {
            std::lock_guard<std::mutex> ws_lock(dev->ws_mu);
            const bool peer = peer_fits(comm, t, count);
            dev->preq = rc_device::PartialReq();
            dev->preq.want = peer;
            bool empty_local = la.size() == 0;
            if (!empty_local) reduce_local(dev, local, t, a, la, axes, dst, ltmp);
            else fill_identity(dev, local, t, static_cast<unsigned char *>(dst) + (size_t)first * es, count);
            const rc_device::PartialReq rq = dev->preq;
            dev->preq = rc_device::PartialReq();
            void *block = static_cast<unsigned char *>(dst) + (size_t)first * es;
            if (peer) {
                if (rq.got) combine_peer(comm, op, t, rq.ptr, rq.S, rq.pitch, count, rq.out, n_reduced_global);
                else combine_peer(comm, op, t, block, 1, 0, count, block, n_reduced_global);
            } else {
                all_reduce_nccl(comm, local, t, block, (size_t)count);
                if (op == RC_MEAN) divide_in_place(dev, t, block, count, n_reduced_global);
            }
        } catch (...) {
            if (tmp) cudaFreeAsync(tmp, dev->stream);
            throw;
        }
        if (tmp) {
            rc_layout lc_c, lt_c;
            to_c(lo, &lc_c);
            to_c(ltmp, &lt_c);
            int st = rc_assign(dev, t, out, &lc_c, t, tmp, &lt_c);
            cudaFreeAsync(tmp, dev->stream);
            if (st != RC_OK) raise((rc_status)st, rc_last_error());
        }
    });
}

}  // extern "C"
