// rc_comm.cu -- the one exchange step of the path: combining per-GPU partials of a sharded reduction
//  with an NCCL all-reduce over NVLink / NVSwitch.  NCCL is bound at run time (dlopen of
// libnccl.so.2, which resolves to the copy torch already loaded when the caller is a torch process) so the
// library has no link-time dependency and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include "rc_device.hpp"
#include "rc_layout.hpp"

struct rc_comm {
    rc_device *dev = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

namespace rc {
namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    });
    RC_CHECK(api.handle && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce,
             RC_ERR_DEVICE, "NCCL (libnccl.so.2) could not be loaded");
    return api;
}

void nccl_check(ncclResult_t r, const char *what) {
    if (r == ncclSuccess) return;
    NcclApi &api = nccl();
    raise(RC_ERR_DEVICE, std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(r) : "NCCL error"));
}

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
    }
    raise(RC_ERR_INVALID_VALUE, "unknown reduction op");
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
        *out = c.release();
    });
}

int rc_comm_destroy(rc_comm *comm) {
    return guard([&] {
        if (!comm) return;
        if (comm->comm) {
            cudaSetDevice(comm->dev->ordinal);
            cudaStreamSynchronize(comm->dev->stream);
            nccl().CommDestroy(comm->comm);
        }
        delete comm;
    });
}

int rc_comm_all_reduce(rc_comm *comm, rc_redop op, rc_dtype t, void *buf, size_t count) {
    return guard([&] {
        RC_CHECK(comm != nullptr, RC_ERR_INVALID_VALUE, "null comm");
        DeviceGuard g(comm->dev);
        if (count == 0) return;
        RC_CHECK(buf != nullptr, RC_ERR_INVALID_VALUE, "null buffer");
        nccl_check(nccl().AllReduce(buf, buf, count, nccl_dtype(t), nccl_op(op), comm->comm, comm->dev->stream),
                   "ncclAllReduce");
    });
}

int rc_reduce_all_sharded(rc_device *dev, rc_comm *comm, rc_redop op, rc_dtype t, const void *a, const rc_layout *la,
                          int64_t n_global, void *host_out) {
    return guard([&] {
        DeviceGuard g(dev);
        RC_CHECK(comm != nullptr && comm->dev == dev, RC_ERR_DEVICE_MISMATCH, "communicator belongs to another device");
        RC_CHECK(host_out != nullptr, RC_ERR_INVALID_VALUE, "null host_out");
        RC_CHECK(op <= RC_MEAN, RC_ERR_UNIMPLEMENTED, "sharded reductions cover sum / prod / max / min / mean");
        void *slot = nullptr;
        cudaError_t e = cudaMallocAsync(&slot, 16, dev->stream);
        if (e != cudaSuccess) raise(RC_ERR_MEMORY, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
        // local partial: mean is carried as a sum and divided by the GLOBAL count at the end
        rc_redop local = (op == RC_MEAN) ? RC_SUM : op;
        int st = rc_reduce_all_device(dev, local, t, a, la, slot);
        if (st == RC_OK) st = rc_comm_all_reduce(comm, local, t, slot, 1);
        unsigned char v[8] = {0};
        if (st == RC_OK) {
            cudaError_t ce = cudaMemcpyAsync(v, slot, dtype_size(t), cudaMemcpyDeviceToHost, dev->stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(dev->stream);
            if (ce != cudaSuccess) { set_last_error(cudaGetErrorString(ce)); st = RC_ERR_DEVICE; }
        }
        cudaFreeAsync(slot, dev->stream);
        if (st != RC_OK) raise((rc_status)st, rc_last_error());
        if (op == RC_MEAN) {
            if (t == RC_F64) { double x; std::memcpy(&x, v, 8); x /= (double)n_global; std::memcpy(v, &x, 8); }
            else if (t == RC_F32) { float x; std::memcpy(&x, v, 4); x /= (float)n_global; std::memcpy(v, &x, 4); }
            else raise(RC_ERR_UNIMPLEMENTED, "mean requires a floating-point dtype");
        }
        std::memcpy(host_out, v, dtype_size(t));
    });
}

}  // extern "C"
