"""Scratch driver for ncu captures of the shape-specialised reduction / unpack kernels: one launch each of
reduce_tiny_kernel ((2^24, 4) sum), reduce_rows_peel_kernel (f32 a[:, 1:-1] sum), reduce_cols_kernel with one row-lane
((4, 2^24) sum axis 0) and unpack_tri_rest_kernel (batch axis fastest)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rstsr_b200 as rt

dev = rt.DeviceCuda(0, rt.ROW_MAJOR)
rng = np.random.default_rng(0)
a = rt.asarray(rng.standard_normal(1 << 26), dev)
for _ in range(2):
    t1 = a.reshape([1 << 24, 4]).sum_axes(-1)
    t2 = a.reshape([4, 1 << 24]).sum_axes(0)
f = rt.asarray(rng.standard_normal(8192 * 8192).astype(np.float32), dev).reshape([8192, 8192])
for _ in range(2):
    t3 = f[:, 1:-1].sum_axes(-1)
batch, m = 256, 256
packed = rt.Tensor(dev.outof_cpu_vec(rng.standard_normal(batch * m * (m + 1) // 2)), rt.Layout((batch, m * (m + 1) // 2), (1, batch)))
for _ in range(2):
    u = packed.unpack_tril("Sy")
dev.synchronize()
print("ok", t1.shape, t2.shape, t3.shape, u.shape)
