#!/usr/bin/env python
"""ew_tile_kernel (LDG/STG + smem tile) vs ew_tile_bulk_kernel (cp.async.bulk / TMA engine) on 8-byte permuted copies.
Run once per setting: RC_TILE_BULK=0 python scripts/probe_tile_bulk.py ; RC_TILE_BULK=1 python scripts/probe_tile_bulk.py
Every result is checked bit for bit against torch."""
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
mode = os.environ.get("RC_TILE_BULK", "default")
if os.environ.get("RC_TILE_WIDE"):
    mode = "wide" + os.environ["RC_TILE_WIDE"]
g = torch.Generator(device="cuda"); g.manual_seed(3)
rows = []


def run(name, shape, perm, dtype=torch.float64, iters=10):
    n = int(np.prod(shape))
    src = torch.rand(n, generator=g, dtype=torch.float64, device="cuda")
    if dtype != torch.float64:
        src = src.to(dtype)
    dst = torch.empty_like(src)
    npdt = {torch.float64: np.float64, torch.int64: np.int64, torch.float32: np.float32}[dtype]
    rs, rd = dev.wrap(src.data_ptr(), n, npdt), dev.wrap(dst.data_ptr(), n, npdt)
    st = [int(np.prod(shape[i + 1:])) for i in range(len(shape))]
    lsrc = Layout(tuple(shape[p] for p in perm), tuple(st[p] for p in perm))
    ldst = Layout.contig(lsrc.shape, rt.ROW_MAJOR)
    l0 = dev.launch_count()
    dev.assign_arbitary(rd, ldst, rs, lsrc)
    launches = dev.launch_count() - l0
    want = src.view(*shape).permute(*perm).contiguous().view(-1)
    ok = bool(torch.equal(dst, want))
    del want
    for _ in range(3):
        dev.assign_arbitary(rd, ldst, rs, lsrc)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dev.assign_arbitary(rd, ldst, rs, lsrc)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    row = {"mode": mode, "case": name, "us": round(us, 1), "gbs": round(2 * n * src.element_size() / us / 1e3, 1), "exact": ok,
           "launches": launches}
    print(json.dumps(row), flush=True)
    rows.append(row)
    assert ok, name


run("cfg2 (1024,1024,512) perm (2,0,1)", (1024, 1024, 512), (2, 0, 1))
run("2-D transpose (16384,16384)", (16384, 16384), (1, 0))
run("batched transpose (64,2048,2048) perm (0,2,1)", (64, 2048, 2048), (0, 2, 1))
run("(512,512,2048) perm (1,2,0)", (512, 512, 2048), (1, 2, 0))
run("4-D (32,64,512,512) perm (1,0,3,2)", (32, 64, 512, 512), (1, 0, 3, 2))
run("small (1024,1024) transpose", (1024, 1024), (1, 0), iters=50)
run("i64 (4096,8192) transpose", (4096, 8192), (1, 0), dtype=torch.int64)
run("f32 2-D transpose (32768,16384)", (32768, 16384), (1, 0), dtype=torch.float32)
run("f32 (1024,1024,1024) perm (2,0,1)", (1024, 1024, 1024), (2, 0, 1), dtype=torch.float32)
run("f32 batched (128,2048,2048) perm (0,2,1)", (128, 2048, 2048), (0, 2, 1), dtype=torch.float32)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"probe_tile_bulk_{mode}.json"), "w"), indent=1)
