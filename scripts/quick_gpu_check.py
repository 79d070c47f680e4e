"""Scratch: first-light check on a B200 -- smoke parity + per-config kernel timings (CUDA events).
Not part of the product or the bench contract; bench.py is the judged harness."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout

PEAK = 6453.1


def timeit(fn, iters=10, warmup=3):
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


def report(name, nbytes, sec, out):
    gbs = nbytes / sec / 1e9
    line = dict(name=name, bytes=nbytes, us=sec * 1e6, gbs=round(gbs, 1), frac_measured=round(gbs / PEAK, 3),
                frac_nominal=round(gbs / 8000, 3))
    print(json.dumps(line), flush=True)
    out.append(line)


def main():
    import __graft_entry__ as g
    g.smoke()
    torch.cuda.set_device(0)
    dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
    out = []
    which = sys.argv[1:] or ["1", "2", "3", "4", "5"]

    def wrap(t):
        return dev.wrap(t.data_ptr(), t.numel(), {torch.float64: np.float64, torch.float32: np.float32}[t.dtype])

    if "1" in which:
        n = 8192
        a = torch.rand(n * n, dtype=torch.float64, device="cuda")
        b = torch.rand(n, dtype=torch.float64, device="cuda")
        c = torch.empty(n * n, dtype=torch.float64, device="cuda")
        la = Layout((n, n), (n, 1)); lb = Layout((n, n), (0, 1)); lc = la
        ra, rb, rc = wrap(a), wrap(b), wrap(c)
        sec = timeit(lambda: dev.op_mutc_refa_refb("add", rc, lc, ra, la, rb, lb))
        report("cfg1 f64 add (8192,8192)+(8192,)", 2 * n * n * 8 + n * 8, sec, out)
        assert torch.equal(c.view(n, n), a.view(n, n) + b)
        sec = timeit(lambda: torch.add(a.view(n, n), b, out=c.view(n, n)))
        report("  torch.add same shapes", 2 * n * n * 8 + n * 8, sec, out)
        sec = timeit(lambda: c.copy_(a))
        report("  torch copy_ 512MiB", 2 * n * n * 8, sec, out)
        del a, b, c

    if "2" in which:
        shp = (1024, 1024, 512)
        N = shp[0] * shp[1] * shp[2]
        src = torch.rand(N, dtype=torch.float64, device="cuda")
        dst = torch.empty(N, dtype=torch.float64, device="cuda")
        rs, rd = wrap(src), wrap(dst)
        lsrc = Layout((512, 1024, 1024), (1, 524288, 512))
        ldst_c = Layout.contig((512, 1024, 1024), rt.ROW_MAJOR)
        ldst_f = Layout.contig((512, 1024, 1024), rt.COL_MAJOR)
        sec = timeit(lambda: dev.assign_arbitary(rd, ldst_c, rs, lsrc), iters=5)
        report("cfg2a f64 permuted copy (2,0,1)->row-major", 2 * N * 8, sec, out)
        ref = src.view(*shp).permute(2, 0, 1).contiguous().view(-1)
        assert torch.equal(dst, ref)
        sec = timeit(lambda: dev.assign_arbitary(rd, ldst_f, rs, lsrc), iters=5)
        report("cfg2b f64 permuted copy (2,0,1)->col-major", 2 * N * 8, sec, out)
        # F-contig of shape (512,1024,1024) == C-contig of reversed shape (1024,1024,512) of permuted (1,0,2)?
        ref = src.view(*shp).permute(2, 0, 1).permute(2, 1, 0).contiguous().view(-1)
        assert torch.equal(dst, ref)
        sec = timeit(lambda: torch.permute(src.view(*shp), (2, 0, 1)).contiguous(), iters=5)
        report("  torch permute(2,0,1).contiguous()", 2 * N * 8, sec, out)
        del src, dst, ref

    if "3" in which:
        n = 16384
        for dt, npdt in ((torch.float32, np.float32), (torch.float64, np.float64)):
            a = torch.rand(n * n, dtype=dt, device="cuda")
            ra = wrap(a)
            la = Layout((n, n), (n, 1))
            es = a.element_size()
            for op in ("sum", "max"):
                for axis in (0, -1):
                    o = torch.empty(n, dtype=dt, device="cuda")
                    ro = wrap(o)
                    lo = Layout((n,), (1,))
                    sec = timeit(lambda: dev.reduce_axes_into(op, ra, la, [axis], ro, lo))
                    report(f"cfg3 {str(dt)[6:]} {op} axis {axis} (16384,16384)", n * n * es + n * es, sec, out)
                    if op == "sum":
                        ref = a.view(n, n).sum(dim=axis)
                        tol = 1e-5 if dt == torch.float32 else 1e-12
                        err = ((o - ref).abs() / ref.abs()).max().item()
                        assert err < tol, (op, axis, err)
                    else:
                        ref = a.view(n, n).max(dim=axis).values
                        assert torch.equal(o, ref)
            sec = timeit(lambda: a.view(n, n).sum(dim=0))
            report(f"  torch {str(dt)[6:]} sum dim0", n * n * es, sec, out)
            sec = timeit(lambda: a.view(n, n).sum(dim=1))
            report(f"  torch {str(dt)[6:]} sum dim1", n * n * es, sec, out)
            del a

    if "4" in which:
        shp = (64, 64, 512, 512)
        N = 64 * 64 * 512 * 512
        a = torch.rand(N, dtype=torch.float64, device="cuda")
        b = torch.rand(N, dtype=torch.float64, device="cuda")
        c = torch.empty(N, dtype=torch.float64, device="cuda")
        ra, rb, rc = wrap(a), wrap(b), wrap(c)
        la = Layout.contig(shp, rt.ROW_MAJOR)
        lb = Layout(shp, (262144, 16777216, 1, 512))
        sec = timeit(lambda: dev.op_mutc_refa_refb("add", rc, la, ra, la, rb, lb), iters=3, warmup=2)
        report("cfg4 f64 c = a + b.transpose(1,0,3,2) (64,64,512,512)", 3 * N * 8, sec, out)
        # check on a slab
        ref = a.view(*shp)[3, 5] + b.view(*shp)[5, 3].t()
        assert torch.equal(c.view(*shp)[3, 5], ref)
        v = torch.rand(512 * 512, dtype=torch.float64, device="cuda")
        rv = wrap(v)
        lv = Layout(shp, (0, 0, 512, 1))
        sec = timeit(lambda: dev.op_mutc_refa_refb("mul", rc, la, ra, la, rv, lv), iters=3, warmup=2)
        report("cfg4b f64 c = a * v (broadcast (0,0,512,1))", 2 * N * 8 + 512 * 512 * 8, sec, out)
        assert torch.equal(c.view(*shp)[7, 9], a.view(*shp)[7, 9] * v.view(512, 512))
        del a, b, c

    if "5" in which:
        N = 1 << 31  # 16 GiB of f64 on one GPU (the per-GPU share of cfg5 at 4 GPUs)
        a = torch.rand(N, dtype=torch.float64, device="cuda")
        ra = wrap(a)
        la = Layout((N,), (1,))
        for op in ("sum", "max"):
            sec = timeit(lambda: dev.reduce_all(op, ra, la), iters=3, warmup=2)
            report(f"cfg5 f64 {op}_all 2^31 elements (16 GiB, incl. D2H of the scalar)", N * 8, sec, out)
        s = dev.reduce_all("sum", ra, la)
        ref = a.sum().item()
        assert abs(s - ref) / ref < 1e-12, (s, ref)
        assert dev.reduce_all("max", ra, la) == a.max().item()
        sec = timeit(lambda: a.sum(), iters=3, warmup=2)
        report("  torch sum 2^31 f64", N * 8, sec, out)
        del a

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "quick_check.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
