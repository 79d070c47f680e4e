"""Scratch driver for ncu captures of the gather kernels: one pack_tril, unpack_tril(Sy), index_select (cols) launch each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rstsr_b200 as rt

dev = rt.DeviceCuda(0, rt.ROW_MAJOR)
batch, m = 64, 1024
full = rt.asarray(np.random.default_rng(0).standard_normal(batch * m * m), dev).reshape([batch, m, m])
for _ in range(2):
    p = full.pack_tril()
    u = p.unpack_tril("Sy")
a = rt.asarray(np.random.default_rng(1).standard_normal(4096 * 4096), dev).reshape([4096, 4096])
idx = np.random.default_rng(2).integers(0, 4096, 4096)
for _ in range(2):
    s = a.index_select(1, idx)
dev.synchronize()
print("ok", float(u.to_numpy()[0, 0, 0]), s.shape)
