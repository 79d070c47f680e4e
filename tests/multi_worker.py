"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun) or a single process (world size 1).

Drives the multi-GPU entry points of the C ABI -- rc_comm_all_reduce, rc_reduce_all_sharded, rc_reduce_axes_sharded --
on shards of ONE seeded array and compares every result with the oracle run on the UNSHARDED array (the semantics
of rstsr-core/src/feature_rayon/auto_impl/reduction.rs:14-63: sum / prod / max / min closures, mean = sum / global n).
Tolerances as the north star states them: integers and max / min bit-exact, f64 sums 1e-12, f32 sums 1e-5 (relative
to the L1 norm).  Prints `MULTI_OK <n checks>` on rank 0 when everything passed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import layout as OL  # noqa: E402

import rstsr_b200 as rt  # noqa: E402
from rstsr_b200 import Layout, shard  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = rt.DeviceCuda(local, rt.ROW_MAJOR)
    uid = [rt.Comm.unique_id() if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(uid, src=0)
    comm = rt.Comm(dev, world, rank, uid[0])
    nranks, myrank, peer = comm.info()
    assert (nranks, myrank) == (world, rank)
    want_peer = os.environ.get("RC_COMM_PEER", "1") != "0"
    if world > 1 and want_peer:
        assert peer, "the peer window could not be set up between the ranks of one box"
    if not want_peer:
        assert not peer
    checks = 0

    def tol(dt, ref_l1):
        dt = np.dtype(dt)
        if dt.kind != "f":
            return 0.0
        return (1e-12 if dt == np.float64 else 1e-5) * ref_l1

    def data(n, dt, seed):
        rng = np.random.default_rng(seed)
        dt = np.dtype(dt)
        if dt.kind == "f":
            return (rng.random(n) * 2.0 - 0.5).astype(dt)
        return rng.integers(-3 if dt.kind == "i" else 0, 4, n).astype(dt)

    # ---- rc_comm_all_reduce: every rank contributes base + rank; expected = fold over ranks in numpy ----
    for dt in (np.float64, np.float32, np.int64, np.int32, np.int16, np.uint16, np.uint8, np.uint64):
        for count in (1, 17, 4096, 32768, 40001, 140001):
            for op in ("sum", "prod", "max", "min"):
                if op == "prod" and count > 4096:
                    continue
                contribs = [data(count, dt, 1000 * r + count) for r in range(world)]
                if op == "prod":  # keep products finite / non-trivial
                    contribs = [np.where(c == 0, 1, c).astype(dt) for c in contribs]
                    if np.dtype(dt).kind == "f":
                        contribs = [(1.0 + 0.001 * c).astype(dt) for c in contribs]
                want = shard.combine_partials(contribs, op)
                buf = dev.outof_cpu_vec(contribs[rank])
                comm.all_reduce(op, buf)
                got = dev.to_cpu_vec(buf)
                if np.dtype(dt).kind == "f" and op in ("sum", "prod"):
                    l1 = np.abs(np.stack(contribs)).sum(axis=0) if op == "sum" else np.abs(want)
                    assert np.all(np.abs(got - want) <= (1e-12 if dt == np.float64 else 1e-5) * l1 + 0), (dt, count, op)
                else:
                    assert np.array_equal(got, want), (dt, count, op)
                checks += 1
                if dist is not None and peer and count * np.dtype(dt).itemsize <= 256 * 1024:
                    # rank-ordered fold: bitwise the same result everywhere
                    import torch
                    t = torch.from_numpy(got.view(np.uint8).copy()).cuda()
                    ref = t.clone()
                    dist.broadcast(ref, src=0)
                    assert torch.equal(t, ref), ("ranks disagree", dt, count, op)

    # ---- rc_reduce_all_sharded vs the oracle on the unsharded array ----
    for dt in (np.float64, np.float32, np.int64, np.int32):
        # the large case takes the two-pass path whose second pass is fused into the combine kernel; f32 stays smaller:
        # the ORACLE's sequential f32 accumulation drifts past 1e-5 of the L1 norm on millions of elements
        for n in (1, 5, 1000, (200003 if dt == np.float32 else 3 * 1024 * 1024 + 7)):
            full = data(n, dt, 77 + n)
            if np.dtype(dt).kind == "f":
                full_prod = (1.0 + 1e-7 * full).astype(dt)
            else:
                full_prod = np.where(full == 0, 1, full).astype(dt)
            for op in ("sum", "max", "min", "mean", "prod"):
                if op == "mean" and np.dtype(dt).kind != "f":
                    continue
                if op == "prod" and n > 1000 and np.dtype(dt).kind == "f":
                    continue  # rounding of a million-factor product depends on the association far beyond 1e-12
                src = full_prod if op == "prod" else full
                lo_, hi_ = shard.shard_bounds(n, world, rank)   # ranks with lo_ == hi_ hold an EMPTY shard
                mine = np.ascontiguousarray(src[lo_:hi_])
                raw = dev.outof_cpu_vec(mine) if mine.size else dev.uninit_impl(dt, 1)
                got = comm.reduce_all_sharded(op, raw, Layout((hi_ - lo_,), (1,)), n)
                want = oracle.reduce_all(op, src, OL.c_contig_layout([n]))
                if np.dtype(dt).kind == "f" and op in ("sum", "mean", "prod"):
                    l1 = float(np.abs(src.astype(np.float64)).sum()) / (n if op == "mean" else 1) if op != "prod" else abs(float(want))
                    assert abs(float(got) - float(want)) <= tol(dt, l1), (dt, n, op, got, want)
                else:
                    assert got == want, (dt, n, op, got, want)
                checks += 1
                again = comm.reduce_all_sharded(op, raw, Layout((hi_ - lo_,), (1,)), n)
                assert again == got or (got != got and again != again), "sharded reduce_all is not run-to-run deterministic"

    # ---- rc_reduce_axes_sharded: the sharded axis is reduced; output full-size on every rank ----
    cases = [((1001, 37), [0], 0), ((64, 515), [0], 0), ((2 * world + 1, 8, 9), [0, 2], 0), ((1, 300), [0], 0),
             ((4099, 1024), [0], 0), ((7, 50000), [0], 0)]
    for shape, axes, shard_axis in cases:
        for dt in (np.float64, np.float32, np.int64):
            n = int(np.prod(shape))
            full = data(n, dt, 5 + n).reshape(shape)
            lfull = OL.c_contig_layout(list(shape))
            lo_, hi_ = shard.shard_bounds(shape[shard_axis], world, rank)
            mine = np.ascontiguousarray(full[lo_:hi_])
            lmine = Layout.contig(mine.shape, rt.ROW_MAJOR)
            raw = dev.outof_cpu_vec(mine.reshape(-1)) if mine.size else dev.uninit_impl(dt, 1)
            n_red = int(np.prod([shape[a] for a in axes]))
            for op in ("sum", "max", "min", "mean", "prod"):
                if op == "mean" and np.dtype(dt).kind != "f":
                    continue
                if op == "prod":
                    continue
                ref, lref = oracle.reduce_axes(op, full.reshape(-1), lfull, axes)
                want = oracle.to_numpy(ref, lref)
                lo = rt.layout_for_reduce(Layout.contig(shape, rt.ROW_MAJOR), axes)
                out = dev.uninit_impl(dt, max(lo.size, 1))
                comm.reduce_axes_sharded(op, raw, lmine, axes, n_red, out, lo)
                got = oracle.to_numpy(dev.to_cpu_vec(out), OL.Layout(lo.shape, lo.stride, lo.offset))
                if np.dtype(dt).kind == "f" and op in ("sum", "mean"):
                    l1 = np.abs(full.astype(np.float64)).sum(axis=tuple(axes)) / (n_red if op == "mean" else 1)
                    assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= tol(dt, 1.0) * l1 + 1e-300), (shape, dt, op)
                else:
                    assert np.array_equal(got, want), (shape, dt, op)
                checks += 1
            # a strided (non-dense) output layout goes through the dense temporary
            if np.dtype(dt) == np.float64 and len(shape) == 2:
                kept = shape[1]
                lo_strided = Layout((kept,), (2,), 1)
                out = dev.zeros_impl(dt, 2 * kept + 2)
                comm.reduce_axes_sharded("sum", raw, lmine, axes, n_red, out, lo_strided)
                got = dev.to_cpu_vec(out)[1:1 + 2 * kept:2]
                want = full.sum(axis=0)
                assert np.all(np.abs(got - want) <= 1e-12 * np.abs(full).sum(axis=0)), (shape, "strided out")
                checks += 1

    # ---- errors are raised on EVERY rank before any exchange (no rank is left waiting) ----
    for op in ("max", "min"):
        try:
            comm.reduce_all_sharded(op, dev.uninit_impl(np.float64, 1), Layout((0,), (1,)), 0)
            raise AssertionError("zero-size max / min must be InvalidValue")
        except rt.RstsrCudaError as e:
            assert e.kind == "InvalidValue"
            checks += 1

    dev.synchronize()
    if dist is not None:
        dist.barrier()
    comm.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        print(f"MULTI_OK {checks} checks, world {world}, peer_window {peer}", flush=True)


if __name__ == "__main__":
    main()
