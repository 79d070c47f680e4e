"""The committed golden vectors (tests/golden/reference_kats.json, made by tests/golden/make_golden.py from the
reference's own test literals) against (1) the oracle, on CPU, and (2) the CUDA path, on the GPU."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import layout as L

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))


def _sliced(order):
    base = L.c_contig_layout([12, 15, 18]) if order == "row" else L.f_contig_layout([12, 15, 18])
    return base.swapaxes(-1, -2).narrow(0, slice(2, -3)).narrow(1, slice(1, -4, 2)).narrow(2, slice(-1, 3, -2))


def test_oracle_matches_golden_reductions():
    a = np.arange(3240, dtype=np.uint64)
    assert oracle.reduce_all("sum", np.arange(24, dtype=np.uint64), L.c_contig_layout([24])) == G["sum_all_arange24"]["expect"]
    assert oracle.reduce_all("sum", a, _sliced("row")) == G["sum_all_sliced_row_major"]["expect"]
    assert oracle.reduce_all("sum", a, _sliced("col")) == G["sum_all_sliced_col_major"]["expect"]
    for key, mk in (("sum_axes_row_major", L.c_contig_layout), ("sum_axes_col_major", L.f_contig_layout)):
        out, lo = oracle.reduce_axes("sum", a, mk([4, 6, 15, 9]).transpose([2, 0, 3, 1]), [0, -2])
        s = oracle.to_numpy(out, lo)
        assert [int(s[tuple(i)]) for i in G[key]["index"]] == G[key]["expect"]
    v = np.array(G["min_4x3"]["data"], dtype=np.int64)
    l = L.c_contig_layout([4, 3])
    assert oracle.to_numpy(*oracle.reduce_axes("min", v, l, [0])).tolist() == G["min_4x3"]["axis0"]
    assert oracle.to_numpy(*oracle.reduce_axes("min", v, l, [1])).tolist() == G["min_4x3"]["axis1"]
    assert oracle.reduce_all("min", v, l) == G["min_4x3"]["all"]
    m = np.arange(24.0)
    for key, mk in (("mean_row_major", L.c_contig_layout), ("mean_col_major", L.f_contig_layout)):
        lm = mk([2, 3, 4])
        assert oracle.to_numpy(*oracle.reduce_axes("mean", m, lm, [0, 2])).tolist() == G[key]["axes_0_2"]
        lv = lm.narrow(0, slice(None, None, -1)).narrow(2, slice(None, None, -2))
        assert oracle.to_numpy(*oracle.reduce_axes("mean", m, lv, [-1, 1])).tolist() == G[key]["flipped_axes_m1_1"]


def test_oracle_matches_golden_layouts_and_elementwise():
    g = G["iter_order_offsets"]
    l = L.Layout(tuple(g["layout"]["shape"]), tuple(g["layout"]["stride"]), g["layout"]["offset"])
    assert list(L.iter_offsets_col_major(L.translate_to_col_major_unary(l, "C"))) == g["C"]
    assert list(L.iter_offsets_col_major(L.translate_to_col_major_unary(l, "K"))) == g["K"]
    assert list(L.bounds_index(l)) == G["bounds_index"]["expect"]
    b = G["broadcast_layout"]
    l1, l2 = L.broadcast_layout(L.c_contig_layout([8, 1, 6, 3, 1]), L.f_contig_layout([7, 1, 3, 5]), "row")
    assert list(l1.shape) == b["shape"] and list(l1.stride) == b["stride1"] and list(l2.stride) == b["stride2"]
    lin = np.linspace
    c, lc = oracle.tensor_binary("add", lin(1, 6, 6), L.c_contig_layout([2, 3]), lin(2, 6, 3), L.c_contig_layout([3]))
    assert oracle.to_numpy(c, lc).reshape(-1).tolist() == G["add_2x3_plus_3"]["expect"]
    c, lc = oracle.tensor_binary("add", lin(1, 6, 6), L.c_contig_layout([1, 2, 3]), lin(1, 10, 10), L.c_contig_layout([5, 1, 2, 1]))
    assert oracle.to_numpy(c, lc).reshape(-1).tolist() == G["add_1x2x3_plus_5x1x2x1"]["expect"]
    r, lr, _ = oracle.tensor_to_contig(np.arange(12), L.c_contig_layout([3, 4]).reverse_axes(), "row")
    assert oracle.to_numpy(r, lr).tolist() == G["to_contig_transposed_3x4"]["expect"]


@pytest.mark.gpu
def test_cuda_path_matches_golden(dev, dev_col):
    import rstsr_b200 as rt
    assert rt.arange(24, dev, dtype=np.uint64).sum_all() == G["sum_all_arange24"]["expect"]
    for d, key_all, key_axes in ((dev, "sum_all_sliced_row_major", "sum_axes_row_major"),
                                 (dev_col, "sum_all_sliced_col_major", "sum_axes_col_major")):
        a = rt.arange(3240, d, dtype=np.uint64)
        assert a.reshape([12, 15, 18]).swapaxes(-1, -2)[2:-3, 1:-4:2, -1:3:-2].sum_all() == G[key_all]["expect"]
        s = a.reshape([4, 6, 15, 9]).transpose([2, 0, 3, 1]).sum_axes([0, -2]).to_numpy()
        assert [int(s[tuple(i)]) for i in G[key_axes]["index"]] == G[key_axes]["expect"]
    v = rt.asarray(np.array(G["min_4x3"]["data"]), dev).reshape([4, 3])
    assert v.min_axes(0).to_vec().tolist() == G["min_4x3"]["axis0"]
    assert v.min_axes(1).to_vec().tolist() == G["min_4x3"]["axis1"]
    assert v.min_all() == G["min_4x3"]["all"]
    for d, key in ((dev, "mean_row_major"), (dev_col, "mean_col_major")):
        m = rt.arange(24, d, dtype=np.float64).reshape([2, 3, 4])
        assert m.mean_axes([0, 2]).to_vec().tolist() == G[key]["axes_0_2"]
        assert m[::-1, :, ::-2].mean_axes([-1, 1]).to_vec().tolist() == G[key]["flipped_axes_m1_1"]
    lin = np.linspace
    a, b = rt.asarray(lin(1, 5, 5), dev), rt.asarray(lin(2, 10, 5), dev)
    assert (a.flip(0) + b).to_numpy().tolist() == G["add_flip_a"]["expect"]
    assert (a + b.flip(0)).to_numpy().tolist() == G["add_flip_b"]["expect"]
    assert (a - b).to_numpy().tolist() == G["sub_5"]["expect"]
    assert (a * b).to_numpy().tolist() == G["mul_5"]["expect"]
    x = rt.asarray(lin(1, 9, 9), dev).reshape([3, 3]) + rt.asarray(lin(2, 18, 9), dev).reshape([3, 3]).reverse_axes()
    assert x.to_numpy().reshape(-1).tolist() == G["add_transposed_3x3"]["expect"]
    t = rt.arange(12, dev).reshape([3, 4]).reverse_axes().to_contig(rt.ROW_MAJOR)
    assert t.to_numpy().tolist() == G["to_contig_transposed_3x4"]["expect"]
    s = rt.arange(24, dev).reshape([4, 6])[::2, ::2]
    g = G["to_contig_sliced_4x6"]
    assert list(s.shape) == g["shape"] and list(s.stride) == g["stride"]
    sc = s.to_contig(rt.ROW_MAJOR)
    assert list(sc.stride) == g["out_stride"] and sc.to_numpy().tolist() == g["expect"]
