"""CPU ORACLE (test infrastructure, NOT product code) -- layout algebra of rstsr-common.

A plain-Python restatement of the reference's host-side layout rules for the DeviceCuda hot path.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this package; the product path (`rstsr_b200/`, `librstsr_cuda.so`) never does.

Every function cites the reference function it follows (paths inside RESTGroup/rstsr v0.7.10).
Pinned by the reference's own known-answer tests, see tests/test_oracle_golden.py:
  rstsr-common/src/layout/rearrangement.rs:465-496 (greedy_layout), broadcast.rs:281+, layoutbase.rs:696+,
  rstsr-core/src/tensor/reduction.rs:417-613, tensor/operators/op_binary_arithmetic.rs:992-1143.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import cmp_to_key
from typing import List, Optional, Sequence, Tuple

ROW_MAJOR = "row"
COL_MAJOR = "col"


class LayoutError(Exception):
    """InvalidLayout / InvalidValue / ValueOutOfRange of rstsr-common/src/error.rs:12-49."""

    def __init__(self, kind: str, msg: str = ""):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind


@dataclass(frozen=True)
class Layout:
    """Layout<IxD> (layoutbase.rs:15-23): shape, element strides (may be 0 / negative), element offset."""

    shape: Tuple[int, ...]
    stride: Tuple[int, ...]
    offset: int = 0

    def __post_init__(self):
        object.__setattr__(self, "shape", tuple(int(x) for x in self.shape))
        object.__setattr__(self, "stride", tuple(int(x) for x in self.stride))
        object.__setattr__(self, "offset", int(self.offset))
        if len(self.shape) != len(self.stride):
            raise LayoutError("InvalidLayout", "shape/stride length mismatch")

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:  # layoutbase.rs:67-69
        n = 1
        for d in self.shape:
            n *= d
        return n

    # PartialEq (layoutbase.rs:547-572)
    def same_as(self, other: "Layout") -> bool:
        if self.ndim != other.ndim or self.offset != other.offset:
            return False
        for d1, d2, s1, s2 in zip(self.shape, other.shape, self.stride, other.stride):
            if d1 != d2:
                return False
            if d1 not in (0, 1) and s1 != s2:
                return False
        return True

    # layoutbase.rs:443-498
    def transpose(self, axes: Sequence[int]) -> "Layout":
        n = self.ndim
        if len(axes) != n:
            raise LayoutError("InvalidLayout", "number of elements in axes should be the same to number of dimensions.")
        ax = normalize_axes(axes, n)
        return Layout(tuple(self.shape[a] for a in ax), tuple(self.stride[a] for a in ax), self.offset)

    def reverse_axes(self) -> "Layout":
        return Layout(self.shape[::-1], self.stride[::-1], self.offset)

    def swapaxes(self, a1: int, a2: int) -> "Layout":
        a1, a2 = check_axis(a1, self.ndim), check_axis(a2, self.ndim)
        ax = list(range(self.ndim))
        ax[a1], ax[a2] = ax[a2], ax[a1]
        return self.transpose(ax)

    # slicing one axis: Layout::dim_narrow (rstsr-common/src/layout/indexer.rs:120-200).  Note: with a
    # negative step an explicit stop of -1 means "down to and including index 0" (unlike NumPy).
    def narrow(self, axis: int, sl: slice) -> "Layout":
        axis = check_axis(axis, self.ndim)
        start, stop, step = sl.start, sl.stop, sl.step
        if start is None and stop is None and step is None:
            return self
        n = self.shape[axis]
        step = 1 if step is None else step
        if step == 0:
            raise LayoutError("InvalidValue", "slice step cannot be zero")
        if n == 0:
            return self
        if step > 0:
            start = 0 if start is None else start
            stop = n if stop is None else stop
            if start < 0:
                start = max(n + start, 0)
            if stop < 0:
                stop = max(n + stop, 0)
            if start > n or start > stop:
                start = stop = 0
            elif stop > n:
                stop = n
            length = max((stop - start + step - 1) // step, 0)
        else:
            start = n - 1 if start is None else start
            stop = -1 if stop is None else stop
            if start < 0:
                start = max(n + start, 0)
            if stop < -1:
                stop = max(n + stop, -1)
            if stop > n - 1 or stop > start:
                start = stop = 0
            elif start > n - 1:
                start = n - 1
            # Rust integer division truncates toward zero
            num = stop - start + step + 1
            q = abs(num) // abs(step)
            length = max(q if (num < 0) == (step < 0) else -q, 0)
        shape = list(self.shape)
        stride = list(self.stride)
        offset = self.offset + stride[axis] * start
        shape[axis] = length
        stride[axis] = stride[axis] * step
        return check_layout(Layout(tuple(shape), tuple(stride), offset))

    def select(self, axis: int, index: int) -> "Layout":
        axis = check_axis(axis, self.ndim)
        n = self.shape[axis]
        if index < 0:
            index += n
        if not 0 <= index < n:
            raise LayoutError("IndexError", "index out of bounds")
        shape = self.shape[:axis] + self.shape[axis + 1:]
        stride = self.stride[:axis] + self.stride[axis + 1:]
        return Layout(shape, stride, self.offset + index * self.stride[axis])

    def insert_axis(self, axis: int) -> "Layout":
        if axis < 0:
            axis += self.ndim + 1
        # new axis of extent 1; stride copied from its right neighbour (or 1)
        st = self.stride[axis] * self.shape[axis] if axis < self.ndim else 1
        return Layout(self.shape[:axis] + (1,) + self.shape[axis:], self.stride[:axis] + (st,) + self.stride[axis:],
                      self.offset)


def check_axis(axis: int, ndim: int) -> int:
    a = axis + ndim if axis < 0 else axis
    if not 0 <= a < ndim:
        raise LayoutError("InvalidValue", f"axis {axis} out of bounds for ndim {ndim}")
    return a


def normalize_axes(axes: Sequence[int], ndim: int) -> List[int]:
    """normalize_axes_index(allow_duplicate=False, sort=False), rstsr-common/src/axis_index.rs:379-414."""
    out = [check_axis(int(a), ndim) for a in axes]
    if len(set(out)) != len(out):
        raise LayoutError("InvalidValue", "Duplicate axes are not allowed.")
    return out


# ---------------------------------------------------------------------------------------------
# contiguity, bounds, validity
# ---------------------------------------------------------------------------------------------
def c_contig_layout(shape: Sequence[int], offset: int = 0) -> Layout:
    """shape.c() (layoutbase.rs:577-581; stride rule shape.rs:85-94)."""
    shape = tuple(int(x) for x in shape)
    stride = [1] * len(shape)
    for i in range(len(shape) - 2, -1, -1):
        stride[i] = stride[i + 1] * max(shape[i + 1], 1)
    return Layout(shape, tuple(stride), offset)


def f_contig_layout(shape: Sequence[int], offset: int = 0) -> Layout:
    shape = tuple(int(x) for x in shape)
    stride = [1] * len(shape)
    for i in range(1, len(shape)):
        stride[i] = stride[i - 1] * max(shape[i - 1], 1)
    return Layout(shape, tuple(stride), offset)


def contig_layout(shape, order, offset=0) -> Layout:
    return c_contig_layout(shape, offset) if order == ROW_MAJOR else f_contig_layout(shape, offset)


def ndim_of_f_contig(l: Layout) -> int:  # layoutbase.rs:152-166
    if l.ndim == 0 or l.size == 0:
        return l.ndim
    acc = 1
    for i, (s, d) in enumerate(zip(l.stride, l.shape)):
        if d != 1 and s != acc:
            return i
        acc *= d
    return l.ndim


def ndim_of_c_contig(l: Layout) -> int:  # layoutbase.rs:172-186
    if l.ndim == 0 or l.size == 0:
        return l.ndim
    acc = 1
    for i, (s, d) in enumerate(zip(reversed(l.stride), reversed(l.shape))):
        if d != 1 and s != acc:
            return i
        acc *= d
    return l.ndim


def f_contig(l: Layout) -> bool:
    return ndim_of_f_contig(l) == l.ndim


def c_contig(l: Layout) -> bool:
    return ndim_of_c_contig(l) == l.ndim


def bounds_index(l: Layout) -> Tuple[int, int]:  # layoutbase.rs:237-262
    if l.ndim == 0:
        return l.offset, l.offset + 1
    lo = hi = l.offset
    for d, s in zip(l.shape, l.stride):
        if d == 0:
            return l.offset, l.offset
        if s > 0:
            hi += s * (d - 1)
        else:
            lo += s * (d - 1)
    if lo < 0:
        raise LayoutError("ValueOutOfRange", "min bound < 0")
    return lo, hi + 1


def check_strides(l: Layout, skip_zero: bool = True) -> None:  # layoutbase.rs:285-320
    if l.size == 0 or l.ndim == 0:
        return
    idx = [k for k in range(l.ndim) if l.shape[k] > 1]
    idx.sort(key=lambda k: abs(l.stride[k]))  # stable
    cum = 0
    for k in idx:
        t = abs(l.stride[k])
        if t == 0 and skip_zero:
            continue
        if not 0 <= cum < t:
            raise LayoutError("InvalidLayout", "stride too small: elements overlap")
        cum += (l.shape[k] - 1) * t


def check_layout(l: Layout) -> Layout:  # Layout::new, layoutbase.rs:396-404
    bounds_index(l)
    check_strides(l, True)
    return l


def size_non_broadcast(l: Layout) -> int:  # broadcast.rs:255-266
    if l.size == 0:
        return 0
    n = 1
    for d, s in zip(l.shape, l.stride):
        if s != 0:
            n *= d
    return n


# ---------------------------------------------------------------------------------------------
# broadcasting (broadcast.rs:21-95, 166-245)
# ---------------------------------------------------------------------------------------------
def broadcast_shape(s1: Sequence[int], s2: Sequence[int], order: str):
    s1, s2 = list(s1), list(s2)
    if order == COL_MAJOR:
        s1.reverse()
        s2.reverse()
    n1, n2 = len(s1), len(s2)
    n = max(n1, n2)
    shape = [0] * n
    tp1 = [None] * n
    tp2 = [None] * n
    for i in range(n - 1, -1, -1):
        i1, i2 = n1 + i - n, n2 + i - n
        d1 = s1[i1] if i1 >= 0 else 1
        d2 = s2[i2] if i2 >= 0 else 1
        if d1 == 1 and d2 == 1:
            tp1[i] = tp2[i] = "preserve"
            shape[i] = 1
        elif d2 == 1:
            tp1[i], tp2[i], shape[i] = "preserve", "upcast", d1
        elif d1 == 1:
            tp1[i], tp2[i], shape[i] = "upcast", "preserve", d2
        else:
            if d1 != d2:
                raise LayoutError("InvalidLayout", "Broadcasting failed.")
            tp1[i] = tp2[i] = "preserve"
            shape[i] = d1
        if i1 < 0:
            tp1[i] = "expand"
        if i2 < 0:
            tp2[i] = "expand"
    if order == COL_MAJOR:
        shape.reverse()
        tp1.reverse()
        tp2.reverse()
    return tuple(shape), tp1, tp2


def _update_layout_by_shape(l: Layout, shape, tp, order) -> Layout:  # broadcast.rs:207-245
    if order == COL_MAJOR:
        r = _update_layout_by_shape(l.reverse_axes(), tuple(reversed(shape)), list(reversed(tp)), ROW_MAJOR)
        return r.reverse_axes()
    n, n_old = len(shape), l.ndim
    stride = [0] * n
    stride[n - n_old:] = list(l.stride)
    for i in range(n):
        if tp[i] in ("expand", "upcast"):
            stride[i] = 0
    return Layout(tuple(shape), tuple(stride), l.offset)


def broadcast_layout(l1: Layout, l2: Layout, order: str) -> Tuple[Layout, Layout]:
    shape, tp1, tp2 = broadcast_shape(l1.shape, l2.shape, order)
    return _update_layout_by_shape(l1, shape, tp1, order), _update_layout_by_shape(l2, shape, tp2, order)


def broadcast_layout_to_first(l1: Layout, l2: Layout, order: str) -> Tuple[Layout, Layout]:
    """broadcast.rs:191-205: the broadcast shape must keep the rank of the first operand."""
    a, b = broadcast_layout(l1, l2, order)
    if a.ndim != l1.ndim:
        raise LayoutError("InvalidLayout", "cannot broadcast to the first operand")
    return a, b


# ---------------------------------------------------------------------------------------------
# iteration-order canonicalisation and output layouts (rearrangement.rs)
# ---------------------------------------------------------------------------------------------
def greedy_layout(l: Layout, keep_shape: bool) -> Tuple[Layout, List[int]]:
    """rearrangement.rs:36-113."""
    if l.size == 0:
        return l, list(range(l.ndim))
    shape, stride, offset = list(l.shape), list(l.stride), l.offset
    if keep_shape:
        for n in range(l.ndim):
            if stride[n] < 0:  # dim_narrow(n, ::-1)
                offset += (shape[n] - 1) * stride[n]
                stride[n] = -stride[n]

    def still(i):
        return shape[i] == 1 or stride[i] == 0

    def cmp(i1, i2):
        b1, b2 = still(i1), still(i2)
        if b1 and b2:
            return -1 if i1 < i2 else (1 if i1 > i2 else 0)
        if b1 and not b2:
            return -1 if keep_shape else 1
        if b2 and not b1:
            return 1 if keep_shape else -1
        a1, a2 = abs(stride[i1]), abs(stride[i2])
        return -1 if a1 < a2 else (1 if a1 > a2 else 0)

    index = sorted(range(l.ndim), key=cmp_to_key(cmp))  # Python's sort is stable like slice::sort_by
    g = Layout(tuple(shape), tuple(stride), offset).transpose(index)
    if not keep_shape:
        gs = [1 if (d == 1 or t == 0) else d for d, t in zip(g.shape, g.stride)]
        gt = [0 if (d == 1 or t == 0) else t for d, t in zip(g.shape, g.stride)]
        g = Layout(tuple(gs), tuple(gt), g.offset)
    return g, index


def reversed_permute(indices: Sequence[int]) -> List[int]:  # rearrangement.rs:116-122
    out = [0] * len(indices)
    for pos, i in enumerate(indices):
        out[i] = pos
    return out


def layout_for_array_copy(l: Layout, order: str = "K", default_order: str = ROW_MAJOR) -> Layout:
    """rearrangement.rs:125-152."""
    if order == "C":
        return c_contig_layout(l.shape)
    if order == "F":
        return f_contig_layout(l.shape)
    if order == "A":
        if c_contig(l):
            return c_contig_layout(l.shape)
        if f_contig(l):
            return f_contig_layout(l.shape)
        return contig_layout(l.shape, default_order)
    if order == "K":
        g, idx = greedy_layout(l, True)
        return f_contig_layout(g.shape).transpose(reversed_permute(idx))
    raise LayoutError("InvalidValue", "Iter order for copy only accepts CFAK.")


def translate_to_col_major_unary(l: Layout, order: str) -> Layout:
    """rearrangement.rs:165-214 (orders C, F, K, G used on the hot path)."""
    if order == "C":
        return l.reverse_axes()
    if order == "F":
        return l
    if order == "K":
        return greedy_layout(l, True)[0]
    if order == "G":
        return greedy_layout(l, False)[0]
    raise LayoutError("InvalidValue", order)


def translate_to_col_major(ls: Sequence[Layout], order: str) -> List[Layout]:
    """rearrangement.rs:232-282."""
    if not ls:
        return []
    if any(l.shape != ls[0].shape for l in ls):
        raise LayoutError("InvalidLayout", "All shape of layout in this function must be the same.")
    if order in ("C", "F"):
        return [translate_to_col_major_unary(l, order) for l in ls]
    if order == "K":
        sizes = [size_non_broadcast(l) for l in ls]
        if max(sizes) == min(sizes):
            k = 0
        else:  # Iterator::max_by_key returns the LAST maximum
            k = max(range(len(ls)), key=lambda i: (sizes[i], i))
        _, perm = greedy_layout(ls[k], True)
        return [l.transpose(perm) for l in ls]
    raise LayoutError("InvalidValue", order)


def translate_to_col_major_with_contig(ls: Sequence[Layout]) -> Tuple[List[Layout], int]:
    """rearrangement.rs:297-322."""
    if not ls:
        return [], 0
    nd = min(ndim_of_f_contig(l) for l in ls)
    if nd == 0:
        return list(ls), 0
    size_contig = 1
    for d in ls[0].shape[:nd]:
        size_contig *= d
    return [Layout(l.shape[nd:], l.stride[nd:], l.offset) for l in ls], size_contig


def get_axes_composition(l: Layout):
    """rearrangement.rs:335-372 -> (size-1 axes, stride-0 axes, contiguous axes, discontiguous axes)."""
    comp_i, comp_b = [], []
    for i in range(l.ndim):
        if l.shape[i] == 1:
            comp_i.append(i)
        elif l.stride[i] == 0:
            comp_b.append(i)
    remain = [i for i in range(l.ndim) if i not in comp_i and i not in comp_b]
    remain.sort(key=lambda i: abs(l.stride[i]))
    comp_c = []
    cur = 1
    for i in remain:
        if l.stride[i] == cur:
            comp_c.append(i)
            cur *= l.shape[i]
    comp_d = [i for i in remain if i not in comp_c]
    return comp_i, comp_b, comp_c, comp_d


def get_layout_for_binary_op(la: Layout, lb: Layout, order: str) -> Layout:
    """rearrangement.rs:394-459."""
    if la.shape != lb.shape:
        raise LayoutError("InvalidLayout", "Shape of two layouts must be the same for this function.")
    ndim = la.ndim
    a1, a0_a, ac_a, _ = get_axes_composition(la)
    _, a0_b, ac_b, _ = get_axes_composition(lb)
    a0_o = [i for i in a0_a if i in a0_b]
    ac_o = []
    for x, y in zip(ac_a, ac_b):
        if x == y:
            ac_o.append(x)
        else:
            break
    ad_o = [i for i in range(ndim) if i not in a0_o and i not in ac_o and i not in a1]
    if order == ROW_MAJOR:
        ad_o.reverse()
    stride = [0] * ndim
    cur = 1
    for i in ac_o + ad_o:
        stride[i] = cur
        cur *= la.shape[i]
    for i in a1:
        s = 1
        if order == ROW_MAJOR:
            for x in stride[i:]:
                if x != 0:
                    s = x
                    break
        else:
            for x in reversed(stride[:i]):
                if x != 0:
                    s = x
                    break
        stride[i] = s
    return Layout(la.shape, tuple(stride), 0)


def dim_split_axes(l: Layout, axes: Sequence[int]) -> Tuple[Layout, Layout]:
    """indexer.rs:453-478 -> (layout of `axes` in the given order, layout of the rest); both keep the offset."""
    ax = normalize_axes(axes, l.ndim)
    rest = [i for i in range(l.ndim) if i not in ax]
    l1 = check_layout(Layout(tuple(l.shape[i] for i in ax), tuple(l.stride[i] for i in ax), l.offset))
    l2 = check_layout(Layout(tuple(l.shape[i] for i in rest), tuple(l.stride[i] for i in rest), l.offset))
    return l1, l2


def layout_for_reduce(l: Layout, axes: Sequence[int]) -> Layout:
    """Output layout of reduce_axes (cpu_rayon/reduction.rs:147-153)."""
    _, lm = dim_split_axes(l, axes)
    return layout_for_array_copy(lm, "K")


# ---------------------------------------------------------------------------------------------
# reshape (reshape.rs)
# ---------------------------------------------------------------------------------------------
def _attempt_nocopy_reshape(old_dims, old_strides, newdims, is_f_order) -> Optional[List[int]]:
    """reshape.rs:8-115 (NumPy's _attempt_nocopy_reshape)."""
    olddims = [d for d in old_dims if d != 1]
    oldstrides = [s for d, s in zip(old_dims, old_strides) if d != 1]
    oldnd, newnd = len(olddims), len(newdims)
    newstrides = [0] * newnd
    oi, oj, ni, nj = 0, 1, 0, 1
    while ni < newnd and oi < oldnd:
        np_, op = newdims[ni], olddims[oi]
        while np_ != op:
            if np_ < op:
                if nj >= newnd:
                    return None
                np_ *= newdims[nj]
                nj += 1
            else:
                if oj >= oldnd:
                    return None
                op *= olddims[oj]
                oj += 1
        for ok in range(oi, oj - 1):
            if is_f_order:
                if oldstrides[ok + 1] != olddims[ok] * oldstrides[ok]:
                    return None
            else:
                if oldstrides[ok] != olddims[ok + 1] * oldstrides[ok + 1]:
                    return None
        if is_f_order:
            newstrides[ni] = oldstrides[oi]
            for nk in range(ni + 1, nj):
                newstrides[nk] = newstrides[nk - 1] * newdims[nk - 1]
        else:
            newstrides[nj - 1] = oldstrides[oj - 1]
            for nk in range(nj - 1, ni, -1):
                newstrides[nk - 1] = newstrides[nk] * newdims[nk]
        ni, nj = nj, nj + 1
        oi, oj = oj, oj + 1
    if ni >= 1:
        last = newstrides[ni - 1]
        if is_f_order:
            last *= newdims[ni - 1]
    else:
        last = 1
    for nk in range(ni, newnd):
        newstrides[nk] = last
    return newstrides


def reshape_substitute_negatives(shape_out: Sequence[int], size_in: int) -> List[int]:  # reshape.rs:120-160
    shape = list(shape_out)
    neg = [i for i, v in enumerate(shape) if v == -1]
    if any(v < -1 for v in shape):
        raise LayoutError("InvalidValue", "Negative index must be -1.")
    if len(neg) > 1:
        raise LayoutError("InvalidValue", "Only one -1 is allowed in shape.")
    if neg:
        rest = 1
        for v in shape:
            if v != -1:
                rest *= v
        if rest == 0 or size_in % rest != 0:
            raise LayoutError("InvalidValue", "Shape '-1' could not be determined")
        shape[neg[0]] = size_in // rest
    return shape


def layout_reshapeable(l: Layout, shape_out: Sequence[int], order: str) -> Optional[Layout]:
    """reshape.rs:170-226: a Layout when the reshape is a view, None when it needs a copy."""
    shape_out = tuple(shape_out)
    size_out = 1
    for v in shape_out:
        size_out *= v
    if size_out != l.size:
        raise LayoutError("InvalidValue", "Size mismatch between input tensor and output tensor.")
    if l.size in (0, 1):
        return check_layout(Layout(shape_out, (1,) * len(shape_out), l.offset))
    if shape_out == l.shape:
        return l
    if order == ROW_MAJOR and c_contig(l):
        return c_contig_layout(shape_out, l.offset)
    if order == COL_MAJOR and f_contig(l):
        return f_contig_layout(shape_out, l.offset)
    st = _attempt_nocopy_reshape(l.shape, l.stride, shape_out, order == COL_MAJOR)
    if st is None:
        return None
    return Layout(shape_out, tuple(st), l.offset)


# ---------------------------------------------------------------------------------------------
# offset iteration (iterator.rs:23-197): IterLayoutColMajor yields offsets with axis 0 fastest
# ---------------------------------------------------------------------------------------------
def iter_offsets_col_major(l: Layout):
    if l.size == 0:
        return
    n = l.ndim
    if n == 0:
        yield l.offset
        return
    idx = [0] * n
    off = l.offset
    while True:
        yield off
        k = 0
        while k < n:
            idx[k] += 1
            off += l.stride[k]
            if idx[k] < l.shape[k]:
                break
            off -= l.stride[k] * l.shape[k]
            idx[k] = 0
            k += 1
        if k == n:
            return
