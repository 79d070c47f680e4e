"""Shard planner for multi-GPU runs (one process per GPU): which axis to split, what each rank owns, and whether
a reduction needs the one exchange step of this path.

  * elementwise / copy / cast: split the OUTERMOST output axis (largest |stride|) into contiguous slabs; broadcast
    operands (stride 0 on that axis) are replicated; no collective.
  * axis reduction: split the outermost NON-REDUCED axis -> no collective (outputs concatenate).  If every
    candidate axis is reduced (or the caller shards a reduced axis), each rank produces a partial output of the
    full output size and the partials are combined with an all-reduce (sum/prod/max/min; mean = sum / global n).
  * full reduction: even contiguous split; all-reduce of one element.

Pure host logic (no CUDA): unit-tested on CPU with world_size-2 gloo process groups.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

from .device import Layout


def shard_bounds(extent: int, nranks: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of range(extent): the first `extent % nranks` ranks get one extra element."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    base, rem = divmod(extent, nranks)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def outermost_axis(layout: Layout, exclude: Sequence[int] = ()) -> Optional[int]:
    """Axis with the largest |stride| among axes of extent > 1 that are not excluded (ties: lowest index)."""
    best, best_stride = None, -1
    ex = {a % layout.ndim for a in exclude} if layout.ndim else set()
    for i, (d, s) in enumerate(zip(layout.shape, layout.stride)):
        if i in ex or d <= 1:
            continue
        if abs(s) > best_stride:
            best, best_stride = i, abs(s)
    return best


def shard_view(layout: Layout, axis: int, nranks: int, rank: int) -> Layout:
    """The rank's slab of `layout` along `axis` as a view of the SAME (global) buffer."""
    start, stop = shard_bounds(layout.shape[axis], nranks, rank)
    shape = list(layout.shape)
    shape[axis] = stop - start
    offset = layout.offset + (start * layout.stride[axis] if stop > start else 0)
    return Layout(tuple(shape), layout.stride, offset)


def local_layout(layout: Layout, axis: int, nranks: int, rank: int) -> Layout:
    """The rank's slab as a layout of its OWN buffer (offset rebased so the slab starts at the rank's element 0).
    Only meaningful when the slab is contiguous in the buffer, i.e. `axis` is the outermost axis."""
    view = shard_view(layout, axis, nranks, rank)
    if layout.stride[axis] == 0:
        return Layout(view.shape, view.stride, layout.offset)  # replicated operand
    lo = view.offset
    for d, s in zip(view.shape, view.stride):
        if d > 0 and s < 0:
            lo += (d - 1) * s
    base = min(lo, view.offset) if view.size else view.offset
    return Layout(view.shape, view.stride, view.offset - base)


@dataclass(frozen=True)
class ReducePlan:
    shard_axis: Optional[int]   # axis of the INPUT that is split across ranks (None: replicate, nranks == 1)
    needs_collective: bool      # True: every rank holds a partial of the full output -> all-reduce
    collective_op: Optional[str]  # "sum" | "prod" | "max" | "min"
    divide_by: Optional[int]    # mean: divide the combined sum by this global count


def plan_reduce(layout: Layout, axes: Optional[Sequence[int]], op: str, nranks: int) -> ReducePlan:
    """axes=None: full reduction.  Prefers an axis that is kept (no collective)."""
    comb = {"sum": "sum", "mean": "sum", "prod": "prod", "max": "max", "min": "min"}[op]
    nd = layout.ndim
    red = list(range(nd)) if axes is None else sorted({a % nd for a in axes})
    n_red = 1
    for a in red:
        n_red *= layout.shape[a]
    div = n_red if op == "mean" else None
    if nranks == 1:
        return ReducePlan(None, False, None, None)
    kept_axis = outermost_axis(layout, exclude=red)
    if kept_axis is not None and layout.shape[kept_axis] >= nranks:
        return ReducePlan(kept_axis, False, None, None)
    ax = outermost_axis(layout)
    if ax is None:  # nothing to split (every extent <= 1): the operand is replicated, a collective would count it nranks times
        return ReducePlan(None, False, None, None)
    return ReducePlan(ax, True, comb, div)


def combine_partials(partials, op: str):
    """Reference combiner used by the CPU tests: what the all-reduce computes, elementwise over numpy arrays."""
    import numpy as np
    fn = {"sum": np.add, "prod": np.multiply, "max": np.fmax, "min": np.fmin}[op]
    out = partials[0].copy()
    for p in partials[1:]:
        out = fn(out, p)
    return out
