#!/usr/bin/env python
"""How far apart may the rows of one tile be?  cfg2-like permuted copies of 4 GiB f64 whose output planes are 8 MB, 512 KB
and 64 KB apart (same tile kernel, same run lengths, same bytes): isolates the page-spread cost of a full-size permutation."""
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
rows = []


def run(name, shape, perm, iters=10):
    n = int(np.prod(shape))
    src = torch.rand(n, dtype=torch.float64, device="cuda")
    dst = torch.empty_like(src)
    rs, rd = dev.wrap(src.data_ptr(), n, np.float64), dev.wrap(dst.data_ptr(), n, np.float64)
    st = [int(np.prod(shape[i + 1:])) for i in range(len(shape))]
    lsrc = Layout(tuple(shape[p] for p in perm), tuple(st[p] for p in perm))
    ldst = Layout.contig(lsrc.shape, rt.ROW_MAJOR)
    dev.assign_arbitary(rd, ldst, rs, lsrc)
    ok = bool(torch.equal(dst, src.view(*shape).permute(*perm).contiguous().view(-1)))
    for _ in range(3):
        dev.assign_arbitary(rd, ldst, rs, lsrc)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dev.assign_arbitary(rd, ldst, rs, lsrc)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    row = {"case": name, "us": round(us, 1), "gbs": round(2 * n * 8 / us / 1e3, 1), "exact": ok}
    print(json.dumps(row), flush=True)
    rows.append(row)
    assert ok


run("cfg2 (1024,1024,512)->(2,0,1): output planes 8 MB apart", (1024, 1024, 512), (2, 0, 1))
run("(4,256,1024,512)->(0,3,1,2): planes 2 MB apart", (4, 256, 1024, 512), (0, 3, 1, 2))
run("(16,64,1024,512)->(0,3,1,2): planes 512 KB apart", (16, 64, 1024, 512), (0, 3, 1, 2))
run("(128,8,1024,512)->(0,3,1,2): planes 64 KB apart", (128, 8, 1024, 512), (0, 3, 1, 2))
run("(1024,1024,512)->(0,2,1): planes 8 KB apart (batched 1024x512 transposes)", (1024, 1024, 512), (0, 2, 1))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "probe_spread.json"), "w"), indent=1)
