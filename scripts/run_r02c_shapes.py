"""ncu driver for the late round-2 reduction kernels: python scripts/run_r02c_shapes.py small3 | small3_kept | arg_rows | half_f32"""
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
shape, axes, op, tdt, ndt = {
    "small3": ((3, 44739242), [0], "sum", torch.float64, np.float64),          # reduce_cols_small_kernel, one-element loads
    "small3_kept": ((174762, 3, 256), [1], "sum", torch.float64, np.float64),  # reduce_cols_small_kernel, 32-byte packs
    "arg_rows": ((8192, 8192), [1], "argmax", torch.float64, np.float64),      # reduce_rows_kernel<PArg>
    "half_f32": ((1342177, 100), [1], "sum", torch.float32, np.float32),       # 16-byte packs
}[case]
n = int(np.prod(shape))
a = torch.rand(n, dtype=tdt, device="cuda")
ra = dev.wrap(a.data_ptr(), n, ndt)
for _ in range(4):
    raw, lo = dev.reduce_axes(op, ra, Layout.contig(shape, rt.ROW_MAJOR), axes)
torch.cuda.synchronize()
