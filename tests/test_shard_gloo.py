"""Host-side multi-GPU logic (rstsr_b200/shard.py) under a real world_size-2 process group (gloo, CPU).
Each rank computes its shard with the CPU oracle (standing in for the device) and the planner decides whether
the exchange step (all-reduce) is needed -- exactly the control flow bench.py runs over NCCL at N > 1."""
import os
import socket

import numpy as np
import pytest

import rstsr_b200 as rt
from rstsr_b200 import shard


def test_shard_bounds_partition():
    for extent in (0, 1, 7, 8, 64, 1000):
        for n in (1, 2, 3, 8):
            spans = [shard.shard_bounds(extent, n, r) for r in range(n)]
            assert spans[0][0] == 0 and spans[-1][1] == extent
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(4, 2, 2)


def test_plans():
    l = rt.Layout((16384, 16384), (16384, 1))
    assert shard.plan_reduce(l, [-1], "sum", 8) == shard.ReducePlan(0, False, None, None)      # rows kept: no collective
    p = shard.plan_reduce(l, [0], "sum", 8)
    assert p.shard_axis == 1 and not p.needs_collective                                        # shard the columns
    p = shard.plan_reduce(l, None, "mean", 8)
    assert p.needs_collective and p.collective_op == "sum" and p.divide_by == 16384 * 16384    # cfg5
    p = shard.plan_reduce(l, None, "max", 4)
    assert p.needs_collective and p.collective_op == "max" and p.divide_by is None
    assert shard.plan_reduce(l, None, "sum", 1).needs_collective is False
    # cfg4: shard on the outermost axis of c; the permuted operand is sharded on the matching (second) buffer axis
    c = rt.Layout.contig((64, 64, 512, 512), rt.ROW_MAJOR)
    assert shard.outermost_axis(c) == 0
    b = rt.Layout((64, 64, 512, 512), (262144, 16777216, 1, 512))
    v = shard.shard_view(b, 0, 8, 3)
    assert v.shape == (8, 64, 512, 512) and v.offset == 24 * 262144


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import oracle
    from oracle import layout as OL
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    full = rng.random(96 * 40)
    lay = rt.Layout((96, 40), (40, 1))
    results = {}
    for op, axes in (("sum", [-1]), ("sum", [0]), ("max", [0]), ("mean", None), ("min", None)):
        plan = shard.plan_reduce(lay, axes, op, world)
        # bench.py shards ROWS for every config; force that to exercise the exchange step for axis-0 reductions
        ax = 0
        view = shard.shard_view(lay, ax, world, rank)
        ol = OL.Layout(view.shape, view.stride, view.offset)
        sharded_axis_reduced = axes is None or (ax in [a % 2 for a in axes])
        local_op = "sum" if op == "mean" else op
        if axes is None:
            part = np.array([oracle.reduce_all(local_op, full, ol)])
        else:
            raw, lo = oracle.reduce_axes(local_op, full, ol, axes)
            part = oracle.to_numpy(raw, lo)
        t = torch.from_numpy(np.ascontiguousarray(part))
        if sharded_axis_reduced:
            red = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[local_op]
            dist.all_reduce(t, op=red)
            got = t.numpy()
            if op == "mean":
                got = got / (96 * 40)
        else:
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            got = np.concatenate([p.numpy() for p in parts])
        results[f"{op}{axes}"] = got
        assert plan.needs_collective == (axes is None), (op, axes, plan)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), **results)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = np.load(os.path.join(str(tmp_path), "out.npz"))
    rng = np.random.default_rng(5)
    full = rng.random(96 * 40).reshape(96, 40)
    assert np.allclose(out["sum[-1]"], full.sum(-1), rtol=1e-13)
    assert np.allclose(out["sum[0]"], full.sum(0), rtol=1e-13)
    assert np.array_equal(out["max[0]"], full.max(0))
    assert np.allclose(out["meanNone"], full.mean(), rtol=1e-13)
    assert out["minNone"][0] == full.min()
