#!/usr/bin/env python
"""Binary add over a sweep of operand layouts (odd extents, slices, broadcasts along one or two axes, short inner
extents, flipped / stepped views): GB/s of algorithmic bytes, next to torch on the same views.  Finds elementwise cliffs."""
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


def lay(t):
    it = t.element_size()
    return Layout(tuple(t.shape), tuple(s for s in t.stride()), 0)


rows = []
for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32)):
    N = 1 << 26
    base_a = torch.rand(N + 4096, dtype=tdt, device="cuda")
    base_b = torch.rand(N + 4096, dtype=tdt, device="cuda")
    out = torch.empty(N + 4096, dtype=tdt, device="cuda")

    def view(buf, shape, strides, off=0):
        return torch.as_strided(buf, shape, strides, off)

    cases = []
    n2 = 8192
    cases.append(("flat odd length", (N - 3,), (1,), (1,), 0, 0))
    cases.append(("flat, operands offset by 1 element", (N - 8,), (1,), (1,), 1, 1))
    cases.append(("flat, only b offset by 1 element", (N - 8,), (1,), (1,), 0, 1))
    cases.append(("rows of 8191 (odd pitch)", (8192, 8191), (8191, 1), (8191, 1), 0, 0))
    cases.append(("a[:, 1:-1] both (pitch 8192)", (8192, 8190), (8192, 1), (8192, 1), 1, 1))
    cases.append(("row broadcast (n,m)+(m,) m=8191", (8192, 8191), (8191, 1), (0, 1), 0, 0))
    cases.append(("col broadcast (n,m)+(n,1)", (8192, 8192), (8192, 1), (1, 0), 0, 0))
    cases.append(("col broadcast odd m=8191", (8192, 8191), (8191, 1), (1, 0), 0, 0))
    cases.append(("3-D (a,b,c)+(a,1,c)", (512, 256, 512), (256 * 512, 512, 1), (512, 0, 1), 0, 0))
    cases.append(("3-D (a,b,c)+(1,b,1)", (512, 256, 512), (256 * 512, 512, 1), (0, 1, 0), 0, 0))
    cases.append(("3-D (a,b,c)+(a,b,1)", (512, 256, 512), (256 * 512, 512, 1), (256, 1, 0), 0, 0))
    cases.append(("short inner 3: (n,3)+(n,3)", (N // 4, 3), (3, 1), (3, 1), 0, 0))
    cases.append(("short inner 3: (n,3)+(3,)", (N // 4, 3), (3, 1), (0, 1), 0, 0))
    cases.append(("short inner 3: (n,3)+(n,1)", (N // 4, 3), (3, 1), (1, 0), 0, 0))
    cases.append(("inner 17 + (17,)", (N // 32, 17), (17, 1), (0, 1), 0, 0))
    cases.append(("stepped a[::2] + b[::2] (useful bytes)", (N // 2,), (2,), (2,), 0, 0))
    for name, shape, sa, sb, oa, ob in cases:
        n = int(np.prod(shape))
        A = view(base_a, shape, sa, oa)
        B = view(base_b, shape, sb, ob)
        C = view(out, shape, tuple(int(np.prod(shape[i + 1:])) for i in range(len(shape))))
        ra, rb, rc = (dev.wrap(t.data_ptr() - o * t.element_size(), N + 4096, ndt) for t, o in ((base_a, 0), (base_b, 0), (out, 0)))
        la, lb = Layout(tuple(shape), tuple(sa), oa), Layout(tuple(shape), tuple(sb), ob)
        lc = Layout.contig(tuple(shape), rt.ROW_MAJOR)
        dev.op_mutc_refa_refb("add", rc, lc, ra, la, rb, lb)
        ok = bool(torch.equal(C, A + B))
        s = timeit(lambda: dev.op_mutc_refa_refb("add", rc, lc, ra, la, rb, lb))
        st = timeit(lambda: torch.add(A, B, out=C))
        it = np.dtype(ndt).itemsize
        nb = (n + n + sum(1 for _ in [0]) * 0) * it  # c + a
        nb += (int(np.prod([d for d, stv in zip(shape, sb) if stv != 0])) if any(stv != 0 for stv in sb) else 1) * it
        row = {"dtype": np.dtype(ndt).name, "case": name, "gbs": round(nb / s / 1e9), "us": round(s * 1e6, 1), "torch_gbs": round(nb / st / 1e9), "ok": ok}
        rows.append(row)
        print(json.dumps(row), flush=True)
        assert ok, row
json.dump(rows, open("gpurun_out/probe_ew_sweep.json", "w"), indent=1)
