#!/usr/bin/env python
"""Generates tests/golden/promotion_table.json from the reference's own promotion table
(rstsr-dtype-traits/src/promotion.rs: the `impl_promotion_asable!(T1, T2, can_cast_self, can_cast_other, Res)` lines,
the bool rule `impl_promotion_bool_T!` and the reflexive impl).  Run in the build container, where /root/reference
exists; the JSON is the committed fixture the tests read (the GPU box has no /root/reference).

    python scripts/gen_promotion_golden.py [/root/reference]
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(os.path.join(ref, "rstsr-dtype-traits", "src", "promotion.rs")).read()

PRIMS = ["i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64", "f32", "f64"]
ALIAS = {"isize": "i64", "usize": "u64"}
table = {}
for t1, t2, res in re.findall(r"impl_promotion_asable!\((\w+), (\w+), \w+, \w+, (\w+)\);", src):
    a, b, r = (ALIAS.get(x, x) for x in (t1, t2, res))
    if a in PRIMS and b in PRIMS:
        prev = table.setdefault(f"{a},{b}", r)
        assert prev == r, (t1, t2, res, prev)
bool_targets = [ALIAS.get(t, t) for t in re.findall(r"impl_promotion_bool_T!\((\w+)\);", src)]
for t in PRIMS:
    assert t in bool_targets, t
    table[f"bool,{t}"] = t
    table[f"{t},bool"] = t
for t in PRIMS + ["bool"]:
    table[f"{t},{t}"] = t
assert len(table) == 121, len(table)
into_float = {t: ("f64" if t[0] in "iu" else t) for t in PRIMS}  # DTypeIntoFloatAPI: promotion.rs:62-118

# ---- half / complex rows of the same file: bool x T (:195-200), complex x primitive (:368-431), primitive x complex
# (:435-512), c32 x c64 (:516-545); no other half rule exists (f16 x f32 is NOT in the reference's table) ----
EXT = {"half::f16": "f16", "half::bf16": "bf16", "c32": "c32", "c64": "c64"}
CPLX = {"f32": "c32", "f64": "c64"}
ext = {}
for t in re.findall(r"impl_promotion_bool_T!\(([\w:]+)\);", src):
    if t in EXT:
        ext[f"bool,{EXT[t]}"] = EXT[t]
        ext[f"{EXT[t]},bool"] = EXT[t]
for kind, tc, tp, res in re.findall(r"impl_promotion_complex_primitive_(cast_self|no_cast_self)!\((\w+), (\w+), \w+, \w+, (\w+)\);", src):
    p = ALIAS.get(tp, tp)
    prev = ext.setdefault(f"{CPLX[tc]},{p}", CPLX[res])
    assert prev == CPLX[res], (tc, tp, res)
for kind, tc, tp, res in re.findall(r"impl_promotion_primitive_complex_(cast_other|nocast_other)!\((\w+), (\w+), \w+, \w+, (\w+)\);", src):
    p = ALIAS.get(tp, tp)
    prev = ext.setdefault(f"{p},{CPLX[tc]}", CPLX[res])
    assert prev == CPLX[res], (tc, tp, res)
assert "impl DTypePromoteAPI<c32> for c64" in src and "impl DTypePromoteAPI<c64> for c32" in src
ext["c32,c64"] = ext["c64,c32"] = "c64"
for t in EXT.values():
    ext[f"{t},{t}"] = t
for c in ("c32", "c64"):  # every primitive pairs with both complex types, in both orders
    for t in PRIMS:
        assert f"{c},{t}" in ext and f"{t},{c}" in ext, (c, t)
into_float_ext = {t: t for t in EXT.values()}  # :85-101
out = {"source": "rstsr-dtype-traits/src/promotion.rs (RESTGroup/rstsr v0.7.10): impl_promotion_asable!, "
                 "impl_promotion_bool_T!, impl<T> DTypePromoteAPI<T> for T; isize = i64, usize = u64",
       "promote": dict(sorted(table.items())), "into_float": into_float,
       "promote_ext": dict(sorted(ext.items())), "into_float_ext": into_float_ext}
path = os.path.join(ROOT, "tests", "golden", "promotion_table.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(path, len(table), len(ext))
