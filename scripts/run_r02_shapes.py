"""ncu driver for the round-2 kernels: one shape per case, a few launches, nothing else on the stream.
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1 -o gpurun_out/r02_<case> python scripts/run_r02_shapes.py <case>
cases: narrow_u8, narrow_i16 (ew_tile_narrow_kernel), outer_f64 (ew_outer_kernel), cast_u8_f32 (ew_kernel, packs),
       tma_cfg2 (ew_tile_tma_kernel, needs RC_TILE_BULK=1), tile_cfg2 (ew_tile_kernel, needs RC_TILE_WIDE=0),
       wide_cfg2 (ew_tile_wide_kernel), short_{f32,u8}_{deint,inter} (ew_tile_short_kernel)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout

case = sys.argv[1]
torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)


def transpose_copy(tdt, ndt, shape, perm):
    n = int(np.prod(shape))
    src = torch.randint(0, 100, (n,), dtype=torch.int32, device="cuda").to(tdt)
    dst = torch.empty_like(src)
    rs, rd = dev.wrap(src.data_ptr(), n, ndt), dev.wrap(dst.data_ptr(), n, ndt)
    st = [int(np.prod(shape[i + 1:])) for i in range(len(shape))]
    lsrc = Layout(tuple(shape[p] for p in perm), tuple(st[p] for p in perm))
    ldst = Layout.contig(lsrc.shape, rt.ROW_MAJOR)
    for _ in range(4):
        dev.assign_arbitary(rd, ldst, rs, lsrc)


if case == "narrow_u8":
    transpose_copy(torch.uint8, np.uint8, (32768, 32768), (1, 0))
elif case == "narrow_i16":
    transpose_copy(torch.int16, np.int16, (32768, 32768), (1, 0))
elif case in ("tma_cfg2", "tile_cfg2"):
    transpose_copy(torch.float64, np.float64, (1024, 1024, 512), (2, 0, 1))
elif case == "outer_f64":
    n = 8192
    col = torch.rand(n, dtype=torch.float64, device="cuda")
    row = torch.rand(n, dtype=torch.float64, device="cuda")
    out = torch.empty(n * n, dtype=torch.float64, device="cuda")
    rc, rr, ro = (dev.wrap(t.data_ptr(), t.numel(), np.float64) for t in (col, row, out))
    for _ in range(4):
        dev.op_mutc_refa_refb("add", ro, Layout((n, n), (n, 1)), rc, Layout((n, n), (1, 0)), rr, Layout((n, n), (0, 1)))
elif case == "cast_u8_f32":
    n = 1 << 27
    a = torch.zeros(n, dtype=torch.uint8, device="cuda")
    b = torch.empty(n, dtype=torch.float32, device="cuda")
    ra, rb = dev.wrap(a.data_ptr(), n, np.uint8), dev.wrap(b.data_ptr(), n, np.float32)
    for _ in range(4):
        dev.assign(rb, Layout((n,), (1,)), ra, Layout((n,), (1,)))
elif case == "wide_cfg2":      # ew_tile_wide_kernel (default for 8-byte permuted copies)
    transpose_copy(torch.float64, np.float64, (1024, 1024, 512), (2, 0, 1))
elif case in ("short_f32_deint", "short_f32_inter", "short_u8_deint", "short_u8_inter"):  # ew_tile_short_kernel
    tdt, ndt = (torch.float32, np.float32) if "f32" in case else (torch.uint8, np.uint8)
    k = 8 if "f32" in case else 3
    n = ((1 << 28) // (k * np.dtype(ndt).itemsize) // 16) * 16 + 16
    transpose_copy(tdt, ndt, (n, k) if "deint" in case else (k, n), (1, 0))
else:
    raise SystemExit(f"unknown case {case}")
torch.cuda.synchronize()
