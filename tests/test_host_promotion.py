"""Operand-type promotion of the mixed-dtype entry points (host logic only, no GPU): rc_dtype_promote and
rc_binop_out_dtype_ex against the reference's own table (tests/golden/promotion_table.json, generated from
rstsr-dtype-traits/src/promotion.rs by scripts/gen_promotion_golden.py) and against the oracle's restatement."""
import json
import os

import numpy as np
import pytest

import rstsr_b200 as rt
from oracle import promotion as PR

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "promotion_table.json")))


def test_oracle_promotion_matches_reference_table():
    assert len(GOLD["promote"]) == 121
    for key, want in GOLD["promote"].items():
        a, b = key.split(",")
        assert PR.promote(a, b) == want, key
    for t, want in GOLD["into_float"].items():
        assert PR.into_float(t) == want


def test_product_promotion_matches_reference_table():
    for key, want in GOLD["promote"].items():
        a, b = key.split(",")
        got = rt.DeviceCuda.promote_types(PR.NP[a], PR.NP[b])
        assert PR.name_of(got) == want, key


def _np_ext(name):
    if name == "bf16":
        return rt.bfloat16
    return PR.NP_EXT.get(name) or PR.NP[name]


def test_half_and_complex_rows_of_the_reference_table():
    """bool x T, complex x primitive, primitive x complex, c32 x c64 (promotion.rs:195-200, :368-545): oracle and product
    against the fixture generated from the reference's macro lines; every other pair with a half type is an error in
    both, as it is a missing impl in the reference."""
    ext = GOLD["promote_ext"]
    assert len(ext) == 54
    names = PR.NAMES + PR.EXT_NAMES
    for a in names:
        for b in names:
            if a not in PR.EXT_NAMES and b not in PR.EXT_NAMES:
                continue
            key = f"{a},{b}"
            usable = not ("bf16" in (a, b) and rt.bfloat16 is None)
            if key in ext:
                assert PR.promote(a, b) == ext[key], key
                if usable:
                    got = rt.DeviceCuda.promote_types(_np_ext(a), _np_ext(b))
                    assert got == np.dtype(_np_ext(ext[key])), key
            else:
                with pytest.raises(TypeError):
                    PR.promote(a, b)
                if usable:
                    with pytest.raises(rt.RstsrCudaError) as e:
                        rt.DeviceCuda.promote_types(_np_ext(a), _np_ext(b))
                    assert e.value.kind == "UnImplemented", key
    for t, want in GOLD["into_float_ext"].items():
        assert PR.into_float(t) == want
    # TOut of the op classes on a promoted complex pair
    D = rt.DeviceCuda
    assert D.binop_out_dtype_ex("add", np.int32, np.complex64) == np.dtype(np.complex128)
    assert D.binop_out_dtype_ex("mul", np.complex64, np.uint16) == np.dtype(np.complex64)
    assert D.binop_out_dtype_ex("eq", np.float64, np.complex64) == np.dtype(np.bool_)
    assert D.binop_out_dtype_ex("div", np.bool_, np.float16) == np.dtype(np.float16)
    with pytest.raises(rt.RstsrCudaError):
        D.binop_out_dtype_ex("pow", np.complex64, np.float32)


OPS = ["add", "sub", "mul", "div", "maximum", "minimum", "floor_divide", "atan2", "copysign", "hypot", "nextafter",
       "logaddexp", "eq", "ne", "lt", "le", "gt", "ge", "pow"]


@pytest.mark.parametrize("op", OPS)
def test_out_dtype_rules(op):
    """TOut per op class (auto_impl/op_ternary_common.rs:6-189) for all 121 operand pairs."""
    for a in PR.NAMES:
        for b in PR.NAMES:
            try:
                _, want = PR.op_types(op, a, b)
            except TypeError:
                with pytest.raises(rt.RstsrCudaError) as e:
                    rt.DeviceCuda.binop_out_dtype_ex(op, PR.NP[a], PR.NP[b])
                assert e.value.kind == "UnImplemented"
                continue
            got = rt.DeviceCuda.binop_out_dtype_ex(op, PR.NP[a], PR.NP[b])
            assert PR.name_of(got) == want, (op, a, b)


def test_powi_restatement_known_values():
    """__powidf2 on exactly representable cases and the negative-exponent reciprocal."""
    a = np.array([2.0, -3.0, 0.5, 10.0])
    assert np.array_equal(PR.powi(a, np.array([10, 3, -2, 0])), np.array([1024.0, -27.0, 4.0, 1.0]))
    a32 = np.array([3.0], dtype=np.float32)
    assert PR.powi(a32, np.array([-1]))[0] == np.float32(1.0) / np.float32(3.0)
    assert np.array_equal(PR.ipow(np.array([3, -2, 7], dtype=np.int8), np.array([5, 7, 3])),
                          np.array([243 - 256, -128, 343 - 256], dtype=np.int8))
