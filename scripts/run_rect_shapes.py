"""ncu driver: one launch of ew_tile_rect_kernel per regime (short Y f64 / f32, short X f64).  Usage:
ncu --set full --clock-control none --import-source on -k regex:ew_tile_rect -o gpurun_out/r01_rect_<case> python scripts/run_rect_shapes.py <case>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt

case = sys.argv[1]
torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
tdt, ndt, k = {"shorty_f64": (torch.float64, np.float64, 17), "shorty_f32": (torch.float32, np.float32, 8),
               "shortx_f64": (torch.float64, np.float64, 17)}[case]
n = (1 << 25) // k
src = torch.rand(k * n, dtype=tdt, device="cuda")
dst = torch.empty(k * n, dtype=tdt, device="cuda")
rs, rd = dev.wrap(src.data_ptr(), k * n, ndt), dev.wrap(dst.data_ptr(), k * n, ndt)
for _ in range(3):
    if case.startswith("shortx"):   # (k, n) viewed transposed -> (n, k) contiguous output: short output rows
        dev.assign(rd, rt.Layout((n, k), (k, 1)), rs, rt.Layout((n, k), (1, n)))
    else:                           # (n, k) viewed transposed -> (k, n): short contiguous runs in the source
        dev.assign(rd, rt.Layout((k, n), (n, 1)), rs, rt.Layout((k, n), (1, k)))
torch.cuda.synchronize()
