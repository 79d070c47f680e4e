#!/usr/bin/env python
"""Sum over one axis for a sweep of shapes (short / long reduced and kept extents, first / middle / last axis): GB/s of the
library and of torch on the same view.  Finds reduction shapes that sit far below the HBM roofline."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)


def timeit(fn, iters=10, warmup=2):
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
N = 1 << 27  # elements
cases = []
for k in (2, 3, 8, 17, 64, 100, 1000, 4096, 65536):
    cases.append(((N // k, k), 1))      # short .. long LAST axis reduced
    cases.append(((k, N // k), 0))      # short .. long FIRST axis reduced
for k in (3, 16, 100):
    cases.append(((256, k, N // (256 * k)), 1))    # middle axis
    cases.append(((N // (256 * k), k, 256), 1))
    cases.append(((N // (64 * k), 64, k), (0, 2)))  # two axes
for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32)):
    item = np.dtype(ndt).itemsize
    for shape, axes in cases:
        axes = (axes,) if isinstance(axes, int) else tuple(axes)
        n = int(np.prod(shape))
        a = torch.rand(n, dtype=tdt, device="cuda")
        ra = dev.wrap(a.data_ptr(), n, ndt)
        la = Layout.contig(shape, rt.ROW_MAJOR)
        oshape = tuple(s for i, s in enumerate(shape) if i not in axes)
        out = torch.empty(max(int(np.prod(oshape)), 1), dtype=tdt, device="cuda")
        ro = dev.wrap(out.data_ptr(), out.numel(), ndt)
        lo = Layout.contig(oshape, rt.ROW_MAJOR)
        s = timeit(lambda: dev.reduce_axes_into("sum", ra, la, list(axes), ro, lo))
        A = a.view(*shape)
        want = A.sum(axes)
        ok = bool(torch.allclose(out.view(want.shape), want, rtol=1e-4 if item == 4 else 1e-11))
        st = timeit(lambda: A.sum(axes))
        nb = n * item + out.numel() * item
        row = {"dtype": np.dtype(ndt).name, "shape": list(shape), "axes": list(axes), "gbs": round(nb / s / 1e9), "us": round(s * 1e6, 1),
               "torch_gbs": round(nb / st / 1e9), "ok": ok}
        rows.append(row)
        print(json.dumps(row), flush=True)
        assert ok, row
        del a, out, want
json.dump(rows, open("gpurun_out/probe_reduce_sweep.json", "w"), indent=1)
worst = sorted(rows, key=lambda r: r["gbs"])[:8]
print("slowest:", json.dumps(worst, indent=1))
