"""ncu driver: argmax over the last / first axis of (8192, 8192) f64 (a few launches).  python scripts/run_arg_shapes.py rows|cols"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
n = 8192
a = torch.rand(n * n, dtype=torch.float64, device="cuda")
ra = dev.wrap(a.data_ptr(), n * n, np.float64)
axis = 1 if sys.argv[1] == "rows" else 0
for _ in range(4):
    raw, lo = dev.reduce_axes("argmax", ra, Layout((n, n), (n, 1)), [axis])
torch.cuda.synchronize()
want = a.view(n, n).argmax(axis)
got = torch.from_numpy(dev.to_cpu_vec(raw).astype(np.int64)).cuda()
assert torch.equal(got, want)
