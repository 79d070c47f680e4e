#!/usr/bin/env python
"""Side-by-side table of gpurun_out/probe_short_<tag>.json files: interleave / de-interleave GB/s per dtype and k."""
import json
import sys

tabs = {t: json.load(open(f"gpurun_out/probe_short_{t}.json")) for t in sys.argv[1:]}
names = list(tabs)
print("dtype k |", " | ".join(names))
for i, r in enumerate(tabs[names[0]]):
    print(r["dtype"], r["k"], "|", " | ".join(f'{tabs[n][i]["interleave_gbs"]:5d}/{tabs[n][i]["deinterleave_gbs"]:5d}' for n in names))
