#!/usr/bin/env python
"""How much does the contiguous RUN LENGTH on the read resp. write side cost on B200?  Plain strided copies (flat vector
kernel, no transpose): rows of `run` f64 taken from / written to a matrix with a 4 KiB row pitch, the other side
contiguous.  Answers whether a wider tile (longer write runs) could lift the permuted-copy kernel."""
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
rows_total = 1 << 21
pitch = 512
big = torch.rand(rows_total * pitch, dtype=torch.float64, device="cuda")      # 8 GiB
rbig = dev.wrap(big.data_ptr(), big.numel(), np.float64)
out = []
for run in (16, 32, 64, 128, 256, 512):
    n = rows_total * run
    small = torch.rand(n, dtype=torch.float64, device="cuda")
    rsmall = dev.wrap(small.data_ptr(), n, np.float64)
    lstr = Layout((rows_total, run), (pitch, 1))
    lcon = Layout((rows_total, run), (run, 1))
    res = {}
    for name, fn in (("read runs", lambda: dev.assign(rsmall, lcon, rbig, lstr)), ("write runs", lambda: dev.assign(rbig, lstr, rsmall, lcon))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        res[name] = round(2 * n * 8 / us / 1e3, 1)
    row = {"run_bytes": run * 8, "strided_read_gbs": res["read runs"], "strided_write_gbs": res["write runs"]}
    print(json.dumps(row), flush=True)
    out.append(row)
    del small
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_runlen.json"), "w"), indent=1)
