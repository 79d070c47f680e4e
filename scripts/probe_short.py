#!/usr/bin/env python
"""(n, k) <-> (k, n) word copies with a packed short axis (ew_tile_short_kernel, rc_tile_short.cuh): GB/s per element size,
k and direction.  n is a multiple of 16 (the vector variant) and NOT a power of two unless POW2=1 (rows a power of two
apart collide in the DRAM address hash).  Knobs: RC_TILE_SHORT=0 (previous kernels), RC_SHORT_TILE_BYTES."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
pow2 = os.environ.get("POW2") == "1"
tag = os.environ.get("TAG", "short")


def timeit(fn, iters=8, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


rows = []
for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32), (torch.int16, np.int16), (torch.uint8, np.uint8)):
    item = np.dtype(ndt).itemsize
    for k in (2, 3, 4, 6, 8, 12, 16, 17, 24, 32, 48, 64):
        n = (1 << 28) // (k * item)
        n = n if pow2 else (n // 16) * 16 + 16
        src = (torch.rand(k * n, device="cuda") * 100).to(tdt)
        dst = torch.empty(k * n, dtype=tdt, device="cuda")
        rs, rd = dev.wrap(src.data_ptr(), k * n, ndt), dev.wrap(dst.data_ptr(), k * n, ndt)
        dev.assign(rd, rt.Layout((n, k), (k, 1)), rs, rt.Layout((n, k), (1, n)))
        ok = bool(torch.equal(dst.view(n, k), src.view(k, n).T))
        a = timeit(lambda: dev.assign(rd, rt.Layout((n, k), (k, 1)), rs, rt.Layout((n, k), (1, n))))
        dev.assign(rd, rt.Layout((k, n), (n, 1)), rs, rt.Layout((k, n), (1, k)))
        ok = ok and bool(torch.equal(dst.view(k, n), src.view(n, k).T))
        b = timeit(lambda: dev.assign(rd, rt.Layout((k, n), (n, 1)), rs, rt.Layout((k, n), (1, k))))
        nb = 2 * k * n * item
        rows.append({"dtype": np.dtype(ndt).name, "k": k, "n": n, "interleave_gbs": round(nb / a / 1e9), "deinterleave_gbs": round(nb / b / 1e9),
                     "exact": ok})
        print(f"{np.dtype(ndt).name} k={k:3d} n={n}: (k,n).T -> (n,k) {nb / a / 1e9:6.0f} GB/s   (n,k).T -> (k,n) {nb / b / 1e9:6.0f} GB/s  exact={ok}",
              flush=True)
        assert ok
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/probe_short_{tag}.json", "w"), indent=1)
