#!/usr/bin/env python
"""Write-only broadcast ops: outer sum / product (n,1) op (1,n) -> (n,n), and (n,n) op (n,1) for reference.
RC_EW_OUTER=0 keeps them on the flat kernel.  GB/s counts the algorithmic bytes (the output, plus streamed inputs)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout

dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
mode = os.environ.get("RC_EW_OUTER", "1")
rows = []


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for tdt, ndt in ((torch.float64, np.float64), (torch.float32, np.float32), (torch.int32, np.int32)):
    n = 8192 if tdt == torch.float64 else 16384
    es = torch.empty(0, dtype=tdt).element_size()
    col = (torch.rand(n, device="cuda") * 100).to(tdt)
    row = (torch.rand(n, device="cuda") * 100).to(tdt)
    full = (torch.rand(n * n, device="cuda") * 100).to(tdt)
    out = torch.empty(n * n, dtype=tdt, device="cuda")
    rc, rr, rf, ro = (dev.wrap(t.data_ptr(), t.numel(), ndt) for t in (col, row, full, out))
    lo = Layout((n, n), (n, 1))
    lcol, lrow = Layout((n, n), (1, 0)), Layout((n, n), (0, 1))
    for name, fn, nbytes, ref in (
        ("outer add col + row", lambda: dev.op_mutc_refa_refb("add", ro, lo, rc, lcol, rr, lrow), n * n * es, lambda: col[:, None] + row[None, :]),
        ("outer sub row - col", lambda: dev.op_mutc_refa_refb("sub", ro, lo, rr, lrow, rc, lcol), n * n * es, lambda: row[None, :] - col[:, None]),
        ("outer mul col * row", lambda: dev.op_mutc_refa_refb("mul", ro, lo, rc, lcol, rr, lrow), n * n * es, lambda: col[:, None] * row[None, :]),
        ("full + col  (n,n)+(n,1)", lambda: dev.op_mutc_refa_refb("add", ro, lo, rf, lo, rc, lcol), 2 * n * n * es, lambda: full.view(n, n) + col[:, None]),
        ("full + row  (n,n)+(n,)", lambda: dev.op_mutc_refa_refb("add", ro, lo, rf, lo, rr, lrow), 2 * n * n * es, lambda: full.view(n, n) + row[None, :]),
    ):
        fn()
        ok = bool(torch.equal(out.view(n, n), ref()))
        us = timeit(fn)
        r = {"outer_kernel": mode, "dtype": np.dtype(ndt).name, "case": name, "us": round(us, 1), "gbs": round(nbytes / us / 1e3, 1), "exact": ok}
        print(json.dumps(r), flush=True)
        rows.append(r)
        assert ok, name
    del col, row, full, out
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"probe_outer_{mode}.json"), "w"), indent=1)
