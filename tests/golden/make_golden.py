"""Writes tests/golden/reference_kats.json: the known-answer vectors of the REFERENCE's own tests for the hot path,
transcribed from the cited files (the reference is Rust and cannot be executed in this image, so the values are
the literals its tests assert; each is also what NumPy gives for the NumPy one-liner the reference test quotes).
Run: python tests/golden/make_golden.py"""
import json
import os

KATS = {
    "sum_all_arange24": {"source": "rstsr-core/src/tensor/reduction.rs:419-421", "expect": 276},
    "sum_all_sliced_row_major": {"source": "rstsr-core/src/tensor/reduction.rs:423-428",
                                 "numpy": "np.arange(3240).reshape(12,15,18).swapaxes(-1,-2)[2:-3,1:-4:2,-1:3:-2].sum()",
                                 "expect": 446586},
    "sum_all_sliced_col_major": {"source": "rstsr-core/src/tensor/reduction.rs:453-466", "expect": 403662},
    "sum_axes_row_major": {"source": "rstsr-core/src/tensor/reduction.rs:488-508",
                           "numpy": "np.arange(3240).reshape(4,6,15,9).transpose(2,0,3,1).sum(axis=(0,-2))",
                           "index": [[0, 1], [1, 2], [3, 5]], "expect": [27270, 154845, 428220]},
    "sum_axes_col_major": {"source": "rstsr-core/src/tensor/reduction.rs:510-530",
                           "index": [[0, 1], [1, 2], [3, 5]], "expect": [217620, 218295, 220185]},
    "min_4x3": {"source": "rstsr-core/src/tensor/reduction.rs:533-556", "data": [8, 4, 2, 9, 3, 7, 2, 8, 1, 6, 10, 5],
                "axis0": [2, 3, 1], "axis1": [2, 3, 1, 5], "all": 1},
    "mean_row_major": {"source": "rstsr-core/src/tensor/reduction.rs:558-586", "all": 11.5, "axes_0_2": [7.5, 11.5, 15.5],
                       "flipped_axes_m1_1": [18.0, 6.0]},
    "mean_col_major": {"source": "rstsr-core/src/tensor/reduction.rs:587-613", "axes_0_2": [9.5, 11.5, 13.5],
                       "flipped_axes_m1_1": [15.0, 14.0]},
    "add_2x3_plus_3": {"source": "rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:1007-1013",
                       "expect": [3., 6., 9., 6., 9., 12.]},
    "add_1x2x3_plus_5x1x2x1": {"source": "rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:1015-1030",
                               "expect": [2., 3., 4., 6., 7., 8., 4., 5., 6., 8., 9., 10., 6., 7., 8., 10., 11., 12., 8., 9.,
                                          10., 12., 13., 14., 10., 11., 12., 14., 15., 16.]},
    "add_transposed_3x3": {"source": "rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:1032-1040",
                           "expect": [3., 10., 17., 8., 15., 22., 13., 20., 27.]},
    "add_flip_a": {"source": "rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:1042-1048", "expect": [7., 8., 9., 10., 11.]},
    "add_flip_b": {"source": "rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:1050-1056", "expect": [11., 10., 9., 8., 7.]},
    "sub_5": {"source": "rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:1146-1154", "expect": [-1., -2., -3., -4., -5.]},
    "mul_5": {"source": "rstsr-core/src/tensor/operators/op_binary_arithmetic.rs:1156-1164", "expect": [2., 8., 18., 32., 50.]},
    "to_contig_transposed_3x4": {"source": "rstsr-core/tests/core_func/manipulation/test_to_contig.rs:56-86",
                                 "expect": [[0, 4, 8], [1, 5, 9], [2, 6, 10], [3, 7, 11]]},
    "to_contig_sliced_4x6": {"source": "rstsr-core/tests/core_func/manipulation/test_to_contig.rs:88-116",
                             "shape": [2, 3], "stride": [12, 2], "out_stride": [3, 1], "expect": [[0, 2, 4], [12, 14, 16]]},
    "iter_order_offsets": {"source": "rstsr-common/src/layout/iterator.rs:1043-1084",
                           "layout": {"shape": [3, 2, 6], "stride": [3, -180, 15], "offset": 782},
                           "C": [782, 797, 812, 827, 842, 857, 602, 617, 632, 647, 662, 677, 785, 800, 815, 830, 845, 860, 605,
                                 620, 635, 650, 665, 680, 788, 803, 818, 833, 848, 863, 608, 623, 638, 653, 668, 683],
                           "K": [602, 605, 608, 617, 620, 623, 632, 635, 638, 647, 650, 653, 662, 665, 668, 677, 680, 683, 782,
                                 785, 788, 797, 800, 803, 812, 815, 818, 827, 830, 833, 842, 845, 848, 857, 860, 863]},
    "bounds_index": {"source": "rstsr-common/src/layout/layoutbase.rs test_bounds_index", "expect": [602, 864]},
    "broadcast_layout": {"source": "rstsr-common/src/layout/broadcast.rs:401-415", "shape": [8, 7, 6, 3, 5],
                         "stride1": [18, 0, 3, 1, 0], "stride2": [0, 1, 0, 7, 21]},
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as f:
        json.dump(KATS, f, indent=1)
    print("wrote", out, len(KATS), "vectors")
