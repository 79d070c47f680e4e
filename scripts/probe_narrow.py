#!/usr/bin/env python
"""Permuted copies of 1- / 2- / 4-byte elements: achieved GB/s (read + write), checked against torch.
RC_TILE_NARROW=0 gives the one-element-per-lane tile kernel for comparison."""
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
mode = os.environ.get("RC_TILE_NARROW", "1")
rows = []
for tdt, ndt in ((torch.uint8, np.uint8), (torch.int16, np.int16), (torch.float32, np.float32)):
    for shape, perm in (((32768, 32768), (1, 0)), ((64, 4096, 4096), (0, 2, 1)), ((1024, 2048, 512), (2, 0, 1))):
        n = int(np.prod(shape))
        src = torch.randint(0, 100, (n,), dtype=torch.int32, device="cuda").to(tdt)
        dst = torch.empty_like(src)
        rs, rd = dev.wrap(src.data_ptr(), n, ndt), dev.wrap(dst.data_ptr(), n, ndt)
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
        for _ in range(10):
            dev.assign_arbitary(rd, ldst, rs, lsrc)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        row = {"narrow_kernel": mode, "dtype": np.dtype(ndt).name, "shape": shape, "perm": perm, "us": round(us, 1),
               "gbs": round(2 * n * src.element_size() / us / 1e3, 1), "exact": ok}
        print(json.dumps(row), flush=True)
        rows.append(row)
        assert ok
        del src, dst
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"probe_narrow_{mode}.json"), "w"), indent=1)
