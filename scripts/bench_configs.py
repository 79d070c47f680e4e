#!/usr/bin/env python
"""Per-config report for ALL five BASELINE.json configurations at 1/2/4/8 GPUs (north star: "throughput is
reported on synthetic tensors of the named shapes at 1, 2, 4 and 8 GPUs, both as GB/s and as a fraction of the
HBM roofline").  bench.py is the judged one-line harness; this script is the detailed table behind it.

    python scripts/bench_configs.py                       # 1 GPU
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_configs.py

STRONG scaling for the sharded configs (cfg4: (64,64,512,512) split on output axis 0; cfg5: 2^33 f64 = 64 GiB
split evenly, capped by what fits), weak for cfg1-3 (each GPU runs the full named shape).
Every number: CUDA events on the launching stream, >= 3 warm-ups, max over ranks, working sets >> L2.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import rstsr_b200 as rt
from rstsr_b200 import Layout, shard

PEAK = 6453.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = rt.DeviceCuda(local, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
    comm = None
    if world > 1:
        uid = [rt.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = rt.Comm(dev, world, rank, uid[0])
    rows = []

    def wrap(t, dt=np.float64):
        return dev.wrap(t.data_ptr(), t.numel(), dt)

    def timeit(fn, iters=10, warmup=3):
        for _ in range(warmup):
            fn()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e-3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def report(name, total_bytes, sec, scaling):
        gbs = total_bytes / sec / 1e9
        row = dict(config=name, n_gpus=world, scaling=scaling, us=round(sec * 1e6, 1), gbs_total=round(gbs, 1),
                   gbs_per_gpu=round(gbs / world, 1), pct_measured=round(gbs / world / PEAK * 100, 1),
                   pct_8TBs=round(gbs / world / 8000 * 100, 1))
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)

    g = torch.Generator(device="cuda")
    g.manual_seed(42 + rank)
    f64 = dict(dtype=torch.float64, device="cuda")

    # ---- cfg1 ----
    n = 8192
    a = torch.rand(n * n, generator=g, **f64); b = torch.rand(n, generator=g, **f64); c = torch.empty(n * n, **f64)
    la, lb = Layout((n, n), (n, 1)), Layout((n, n), (0, 1))
    ra, rb, rc = wrap(a), wrap(b), wrap(c)
    sec = timeit(lambda: dev.op_mutc_refa_refb("add", rc, la, ra, la, rb, lb))
    report("cfg1 f64 add (8192,8192)+(8192,)", world * (2 * n * n * 8 + n * 8), sec, "weak")
    del a, b, c

    # ---- cfg2 ----
    shp = (1024, 1024, 512)
    N = shp[0] * shp[1] * shp[2]
    src = torch.rand(N, generator=g, **f64); dst = torch.empty(N, **f64)
    rs, rd = wrap(src), wrap(dst)
    lsrc = Layout((512, 1024, 1024), (1, 524288, 512))
    for order, nm in ((rt.ROW_MAJOR, "RowMajor"), (rt.COL_MAJOR, "ColMajor")):
        ldst = Layout.contig((512, 1024, 1024), order)
        sec = timeit(lambda: dev.assign_arbitary(rd, ldst, rs, lsrc), iters=5)
        report(f"cfg2 f64 (1024,1024,512).transpose(2,0,1).to_contig({nm})", world * 2 * N * 8, sec, "weak")
    del src, dst

    # ---- cfg3 ----
    n = 16384
    for dt, npdt in ((torch.float32, np.float32), (torch.float64, np.float64)):
        m = torch.rand(n * n, generator=g, dtype=dt, device="cuda")
        o = torch.empty(n, dtype=dt, device="cuda")
        rm, ro = wrap(m, npdt), wrap(o, npdt)
        lm, lo = Layout((n, n), (n, 1)), Layout((n,), (1,))
        es = m.element_size()
        for op in ("sum", "max"):
            for axis in (0, -1):
                sec = timeit(lambda: dev.reduce_axes_into(op, rm, lm, [axis], ro, lo))
                report(f"cfg3 {str(dt)[6:]} {op} axis {axis} (16384,16384)", world * (n * n * es + n * es), sec, "weak")
        del m, o

    # ---- cfg4: strong scaling, sharded on output axis 0 ----
    full = Layout.contig((64, 64, 512, 512), rt.ROW_MAJOR)
    i0, i1 = shard.shard_bounds(64, world, rank)
    ni = i1 - i0
    loc = ni * 64 * 512 * 512
    a = torch.rand(loc, generator=g, **f64); bl = torch.rand(loc, generator=g, **f64); c = torch.empty(loc, **f64)
    ra, rb, rc = wrap(a), wrap(bl), wrap(c)
    lc = Layout.contig((ni, 64, 512, 512), rt.ROW_MAJOR)
    # b_r: local C-contiguous (64, ni, 512, 512) buffer viewed permuted (1,0,3,2) -> shape (ni,64,512,512)
    lbp = Layout((ni, 64, 512, 512), (262144, ni * 262144, 1, 512))
    sec = timeit(lambda: dev.op_mutc_refa_refb("add", rc, lc, ra, lc, rb, lbp), iters=5)
    report("cfg4 f64 c = a + b.transpose(1,0,3,2) (64,64,512,512) sharded on axis 0", 3 * 64 * 64 * 512 * 512 * 8, sec,
           "strong")
    v = torch.rand(512 * 512, generator=g, **f64)
    rv = wrap(v)
    lv = Layout((ni, 64, 512, 512), (0, 0, 512, 1))
    sec = timeit(lambda: dev.op_mutc_refa_refb("mul", rc, lc, ra, lc, rv, lv), iters=5)
    report("cfg4b f64 c = a * v, v broadcast (0,0,512,1)", 2 * 64 * 64 * 512 * 512 * 8 + world * 512 * 512 * 8, sec, "strong")
    del a, bl, c

    # ---- cfg5: 2^33 f64 (64 GiB) split evenly; a single GPU holds at most 2^34 bytes... use what the config says ----
    total = 1 << 33
    per = total // world
    x = torch.rand(per, generator=g, **f64)
    rx = wrap(x)
    lx = Layout((per,), (1,))
    for op in ("sum", "max"):
        if comm is None:
            sec = timeit(lambda: dev.reduce_all(op, rx, lx), iters=5)
        else:
            sec = timeit(lambda: comm.reduce_all_sharded(op, rx, lx, total), iters=5)
        report(f"cfg5 f64 {op}_all 2^33 elements (64 GiB) incl. all-reduce + D2H of the scalar", total * 8, sec, "strong")
    if comm is not None:
        s = comm.reduce_all_sharded("sum", rx, lx, total)
        ref = torch.tensor([x.sum().item()], dtype=torch.float64, device="cuda")
        dist.all_reduce(ref)
        assert abs(s - ref.item()) / ref.item() < 1e-12
        m = comm.reduce_all_sharded("max", rx, lx, total)
        refm = torch.tensor([x.max().item()], dtype=torch.float64, device="cuda")
        dist.all_reduce(refm, op=dist.ReduceOp.MAX)
        assert m == refm.item()
    del x

    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"configs_n{world}.json"), "w") as f:
            json.dump(rows, f, indent=1)
    if dist is not None:
        comm.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
