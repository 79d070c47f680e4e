"""ncu driver: the outer sum (n,1) + (1,n) -> (n,n) f64 (write-only op of two broadcast operands, flat vector kernel).
ncu --set full --clock-control none --import-source on -k regex:ew_kernel -s 2 -c 1 -o gpurun_out/r01_outer python scripts/run_outer.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
n = 8192
col = torch.rand(n, dtype=torch.float64, device="cuda")
row = torch.rand(n, dtype=torch.float64, device="cuda")
out = torch.empty(n * n, dtype=torch.float64, device="cuda")
rc, rr, ro = (dev.wrap(t.data_ptr(), t.numel(), np.float64) for t in (col, row, out))
lo = rt.Layout((n, n), (n, 1))
for _ in range(3):
    dev.op_mutc_refa_refb("add", ro, lo, rc, rt.Layout((n, n), (1, 0)), rr, rt.Layout((n, n), (0, 1)))
torch.cuda.synchronize()
