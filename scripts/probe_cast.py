"""Scratch probe: achieved bandwidth of casting copies, fills and a few strided patterns (not product)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)


def timeit(fn, iters=20, warmup=3):
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


def rep(name, nbytes, fn):
    s = timeit(fn)
    print(f"{name:58s} {nbytes / s / 1e9:8.1f} GB/s  {s * 1e6:8.1f} us", flush=True)


N = 1 << 27
flat = Layout((N,), (1,))
bufs = {}
for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32), (torch.int32, np.int32), (torch.uint8, np.uint8),
                 (torch.int64, np.int64), (torch.int16, np.int16)):
    t = torch.zeros(N, dtype=tdt, device="cuda")
    bufs[np.dtype(ndt).name] = (t, dev.wrap(t.data_ptr(), N, ndt))
for src, dst in (("float32", "float64"), ("float64", "float32"), ("int32", "float64"), ("uint8", "float32"), ("float64", "int64"),
                 ("int16", "int32"), ("float32", "float32")):
    ts, rs = bufs[src]
    td, rd = bufs[dst]
    nb = N * (np.dtype(src).itemsize + np.dtype(dst).itemsize)
    rep(f"assign cast {src} -> {dst}", nb, lambda: dev.assign(rd, flat, rs, flat))
    rep(f"   torch copy_ {src} -> {dst}", nb, lambda: td.copy_(ts))
t64, r64 = bufs["float64"]
rep("fill f64", N * 8, lambda: dev.fill(r64, flat, 1.5))
rep("   torch fill_", N * 8, lambda: t64.fill_(1.5))
half = Layout((N // 2,), (2,))
rep("assign strided [::2] f64 -> contiguous (useful bytes)", N // 2 * 16, lambda: dev.assign(bufs["int64"][1].__class__(dev, bufs["int64"][1].ptr, N, np.float64, owned=False), Layout((N // 2,), (1,)), r64, half))
rep("neg f64 (unary out of place)", N * 16, lambda: dev.unary_muta_refb("neg", bufs["int64"][1].__class__(dev, bufs["int64"][1].ptr, N, np.float64, owned=False), flat, r64, flat))
rep("exp f64", N * 16, lambda: dev.unary_muta_refb("exp", bufs["int64"][1].__class__(dev, bufs["int64"][1].ptr, N, np.float64, owned=False), flat, r64, flat))
rep("   torch exp f64", N * 16, lambda: torch.exp(t64))
