#!/usr/bin/env python
"""Why is cfg2 slower inside the bench step than alone?  Times the permuted copy (a) back to back with itself,
(b) after each of the other step kernels, (c) after an idle gap, (d) after an L2-sized dirty write, with CUDA events
around the copy only.  Prints one JSON line per case; writes gpurun_out/probe_instep.json."""
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
f64 = dict(dtype=torch.float64, device="cuda")
N1, SHP2, N3 = 8192, (1024, 1024, 512), 16384
g = torch.Generator(device="cuda"); g.manual_seed(1)
a1 = torch.rand(N1 * N1, generator=g, **f64); b1 = torch.rand(N1, generator=g, **f64); c1 = torch.empty(N1 * N1, **f64)
src2 = torch.rand(SHP2[0] * SHP2[1] * SHP2[2], generator=g, **f64); dst2 = torch.empty_like(src2)
m3 = torch.rand(N3 * N3, generator=g, **f64); o3 = torch.empty(N3, **f64)
scratch = torch.empty(64 << 20, dtype=torch.float32, device="cuda")  # 256 MiB
w = lambda t: dev.wrap(t.data_ptr(), t.numel(), np.float64)
ra1, rb1, rc1, rs2, rd2, rm3, ro3 = map(w, (a1, b1, c1, src2, dst2, m3, o3))
la1, lb1 = Layout((N1, N1), (N1, 1)), Layout((N1, N1), (0, 1))
lsrc2 = Layout((SHP2[2], SHP2[0], SHP2[1]), (1, SHP2[1] * SHP2[2], SHP2[2]))
ldst2 = Layout.contig(lsrc2.shape, rt.ROW_MAJOR)
ldst2f = Layout.contig(lsrc2.shape, rt.COL_MAJOR)
lm3, lo3 = Layout((N3, N3), (N3, 1)), Layout((N3,), (1,))

cfg1 = lambda: dev.op_mutc_refa_refb("add", rc1, la1, ra1, la1, rb1, lb1)
cfg2 = lambda: dev.assign_arbitary(rd2, ldst2, rs2, lsrc2)
cfg2f = lambda: dev.assign_arbitary(rd2, ldst2f, rs2, lsrc2)
rows = lambda: dev.reduce_axes_into("sum", rm3, lm3, [-1], ro3, lo3)
cols = lambda: dev.reduce_axes_into("sum", rm3, lm3, [0], ro3, lo3)
dirty = lambda: scratch.fill_(1.0)
idle = lambda: torch.cuda._sleep(2_000_000)  # ~1 ms of spinning, no memory traffic


def measure(name, before, target, after=None, reps=10):
    for _ in range(3):
        for f in before:
            f()
        target()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        for f in before:
            f()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); target(); e1.record()
        if after:
            after()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    row = {"case": name, "us_median": round(ts[len(ts) // 2], 1), "us_min": round(ts[0], 1), "us_max": round(ts[-1], 1)}
    print(json.dumps(row), flush=True)
    return row


out = []
out.append(measure("cfg2 after cfg2 (steady state)", [cfg2], cfg2))
out.append(measure("cfg2 after idle spin", [idle], cfg2))
out.append(measure("cfg2 after cfg1 (as in the step)", [cfg1], cfg2))
out.append(measure("cfg2 after cfg3 cols (read-only predecessor)", [cols], cfg2))
out.append(measure("cfg2 after cfg3 rows", [rows], cfg2))
out.append(measure("cfg2 after 256 MiB fill (dirty L2)", [dirty], cfg2))
out.append(measure("cfg2 ColMajor after cfg2 ColMajor", [cfg2f], cfg2f))
out.append(measure("cfg2 ColMajor after cfg1", [cfg1], cfg2f))
out.append(measure("cfg3 rows after cfg3 rows", [rows], rows))
out.append(measure("cfg3 rows after cfg2 (as in the step)", [cfg2], rows))
out.append(measure("cfg3 rows after idle", [idle], rows))
out.append(measure("cfg1 after cfg1", [cfg1], cfg1))
out.append(measure("cfg1 after cfg3 cols (as in the step)", [cols], cfg1))
# whole step with and without per-op events
def step():
    cfg1(); cfg2(); rows(); cols()
for _ in range(3):
    step()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record(); torch.cuda.synchronize()
row = {"case": "whole step, no per-op events, 20 steps", "us_per_step": round(e0.elapsed_time(e1) * 1e3 / 20, 1)}
print(json.dumps(row)); out.append(row)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_instep.json"), "w"), indent=1)
