// rc_device.hpp -- the device handle behind `rc_device *` and CUDA error plumbing.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "rc_common.hpp"

struct rc_device {
    int ordinal = 0;
    rc_order order = RC_ROW_MAJOR;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    std::atomic<uint64_t> launches{0};
    // reduction workspace (partials of two-pass grid reductions), grown lazily
    std::mutex ws_mu;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    // Sharded reductions (rc_comm.cu) ask the two-pass launchers to stop after the FIRST pass: the fold over the
    // local partial states is fused into the cross-GPU combine kernel.  Set / read under ws_mu.
    struct PartialReq {
        bool want = false;      // in: stop after the first pass when the policy's state is its element type
        bool got = false;       // out: `ptr` holds states[S][pitch]; `out` was not written
        const void *ptr = nullptr;
        int64_t S = 0, pitch = 0;
        void *out = nullptr;    // typed output pointer (base offset applied), contiguous in canonical order
    } preq;
    // 64-byte pinned + mapped host slot: kernels of `*_all` reductions write their scalar straight into host
    // memory, so the call costs launch + stream sync (no allocation, no D2H copy).  Held under slot_mu from the
    // launch until the host has read the value.
    std::mutex slot_mu;
    void *slot_host = nullptr;
    void *slot_dev = nullptr;
    // pinned staging ring for small host arguments that must reach the device (index lists of rc_index_select):
    // cudaMemcpyAsync from PAGEABLE memory synchronises the stream first, which would serialise back-to-back calls.
    static constexpr int STAGE_SLOTS = 4;
    static constexpr size_t STAGE_BYTES = 1u << 20;
    std::mutex stage_mu;
    void *stage_buf[STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    int stage_next = 0;
};

namespace rc {

#define RC_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            ::rc::raise(RC_ERR_DEVICE, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

// cudaSetDevice for the duration of one entry point
struct DeviceGuard {
    explicit DeviceGuard(const rc_device *d) {
        RC_CHECK(d != nullptr, RC_ERR_INVALID_VALUE, "null device handle");
        RC_CUDA(cudaSetDevice(d->ordinal));
    }
};

inline void after_launch(rc_device *d, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) raise(RC_ERR_DEVICE, std::string(what) + " launch failed: " + cudaGetErrorString(e));
    d->launches.fetch_add(1, std::memory_order_relaxed);
}

void *workspace(rc_device *d, size_t nbytes);  // stream-ordered scratch, valid until the next call
// H2D copy of a small host array without blocking on the stream: through the handle's pinned ring when it fits
// (<= STAGE_BYTES), else a plain cudaMemcpyAsync.  `src` may be freed as soon as the call returns.
void upload_small(rc_device *d, void *dst_dev, const void *src, size_t nbytes);
void *scalar_slot(rc_device *d, void **host);  // device view of the handle's mapped host slot (caller holds slot_mu)

}  // namespace rc
