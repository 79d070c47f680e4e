"""Scratch probe: achieved HBM bandwidth of index_select / pack_tri / unpack_tri next to torch (not product)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def rep(name, nbytes, fn):
    s = timeit(fn)
    print(f"{name:62s} {nbytes / s / 1e9:8.1f} GB/s  {s * 1e6:8.1f} us", flush=True)


n = 8192
N = n * n
a = torch.rand(N, dtype=torch.float64, device="cuda")
c = torch.empty(N, dtype=torch.float64, device="cuda")
ra, rc = dev.wrap(a.data_ptr(), N, np.float64), dev.wrap(c.data_ptr(), N, np.float64)
full = Layout((n, n), (n, 1))
rng = np.random.default_rng(0)
idx = rng.integers(0, n, n).astype(np.int64)
tidx = torch.from_numpy(idx).cuda()
A = a.view(n, n)
rep("index_select rows (8192,8192) f64, 8192 random rows", 2 * N * 8, lambda: dev.index_select(rc, full, ra, full, 0, idx))
rep("index_select cols (8192,8192) f64, 8192 random cols", 2 * N * 8, lambda: dev.index_select(rc, full, ra, full, 1, idx))
rep("torch.index_select rows", 2 * N * 8, lambda: torch.index_select(A, 0, tidx))
rep("torch.index_select cols", 2 * N * 8, lambda: torch.index_select(A, 1, tidx))

# long rows (256 KiB each): clustered selections take the windowed kernel, scattered ones the direct kernel
m_rows, n_src = 2048, 32768
Nl = m_rows * n_src
al = torch.rand(Nl, dtype=torch.float64, device="cuda")
Al = al.view(m_rows, n_src)
ra_l = dev.wrap(al.data_ptr(), Nl, np.float64)
for name, ix in (("every other column", np.arange(0, n_src, 2)), ("random 70 % mask", np.nonzero(rng.random(n_src) < 0.7)[0]),
                 ("scattered random", rng.integers(0, n_src, n_src // 2))):
    ix = ix.astype(np.int64)
    cl = torch.empty(m_rows * ix.size, dtype=torch.float64, device="cuda")
    rc_l = dev.wrap(cl.data_ptr(), cl.numel(), np.float64)
    lsrc, ldst = Layout((m_rows, n_src), (n_src, 1)), Layout((m_rows, ix.size), (ix.size, 1))
    tix = torch.from_numpy(ix).cuda()
    dev.index_select(rc_l, ldst, ra_l, lsrc, 1, ix)
    assert torch.equal(cl.view(m_rows, ix.size), torch.index_select(Al, 1, tix))
    rep(f"index_select cols of (2048,32768) f64: {name} (output bytes x 2)", 2 * cl.numel() * 8, lambda: dev.index_select(rc_l, ldst, ra_l, lsrc, 1, ix))
    rep("   torch.index_select (indices on the device)", 2 * cl.numel() * 8, lambda: torch.index_select(Al, 1, tix))
    del cl
del al

for batch, m in ((64, 1024), (4, 4096)):
    tp = m * (m + 1) // 2
    nf, npk = batch * m * m, batch * tp
    f = torch.rand(nf, dtype=torch.float64, device="cuda")
    p = torch.empty(npk, dtype=torch.float64, device="cuda")
    rf, rp = dev.wrap(f.data_ptr(), nf, np.float64), dev.wrap(p.data_ptr(), npk, np.float64)
    lf = Layout((batch, m, m), (m * m, m, 1))
    lp = Layout((batch, tp), (tp, 1))
    rep(f"pack_tril ({batch},{m},{m}) f64", 2 * npk * 8, lambda: dev.pack_tri(rp, lp, rf, lf, "L"))
    rep(f"pack_triu ({batch},{m},{m}) f64", 2 * npk * 8, lambda: dev.pack_tri(rp, lp, rf, lf, "U"))
    rep(f"unpack_tril Sy ({batch},{m},{m}) f64 [read tp + write n^2]", (npk + nf) * 8, lambda: dev.unpack_tri(rf, lf, rp, lp, "L", "Sy"))
    rep(f"unpack_triu Ay ({batch},{m},{m}) f64", (npk + nf) * 8, lambda: dev.unpack_tri(rf, lf, rp, lp, "U", "Ay"))
    rep(f"unpack_tril N ({batch},{m},{m}) f64 [read tp + write tp]", 2 * npk * 8, lambda: dev.unpack_tri(rf, lf, rp, lp, "L", "N"))
    F = f.view(batch, m, m)
    ti = torch.tril_indices(m, m, device="cuda")
    rep("torch F[:, i, j] gather by tril_indices", 2 * npk * 8, lambda: F[:, ti[0], ti[1]])
    del f, p

# batch axis fastest in memory (the reference's own rayon test: [4, 64, 256, 256].f() on a row-major device)
batch, m = 256, 256
tp = m * (m + 1) // 2
nf, npk = batch * m * m, batch * tp
f = torch.rand(nf, dtype=torch.float64, device="cuda")
p = torch.empty(npk, dtype=torch.float64, device="cuda")
rf, rp = dev.wrap(f.data_ptr(), nf, np.float64), dev.wrap(p.data_ptr(), npk, np.float64)
lf = Layout((batch, m, m), (1, batch, batch * m))
lp = Layout((batch, tp), (1, batch))
rep(f"F-layout pack_tril ({batch},{m},{m}) f64", 2 * npk * 8, lambda: dev.pack_tri(rp, lp, rf, lf, "L"))
rep(f"F-layout unpack_tril Sy ({batch},{m},{m}) f64", (npk + nf) * 8, lambda: dev.unpack_tri(rf, lf, rp, lp, "L", "Sy"))
rep(f"F-layout unpack_triu Ah ({batch},{m},{m}) f64", (npk + nf) * 8, lambda: dev.unpack_tri(rf, lf, rp, lp, "U", "Ah"))
