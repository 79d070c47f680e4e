"""Aggregate gpurun_out/r01_launches.csv (ncu --metrics gpu__time_duration.sum launch list) into
profiles/r01_launches_summary.md, copy the csv to profiles/, and refresh profiles/traffic.json from
profiles/r01_ncu_full_summary.json.  Read here, no GPU.  Usage: python scripts/launch_summary.py [round]"""
import csv
import io
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RND = sys.argv[1] if len(sys.argv) > 1 else "r01"


def short(name):
    name = re.sub(r"\(.*$", "", name)           # drop the parameter list
    name = name.replace("<unnamed>::", "").replace("rc::", "")
    return name[:150]


def main():
    src = os.path.join(ROOT, "gpurun_out", f"{RND}_launches.csv")
    text = open(src).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = {}
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        k = short(r["Kernel Name"])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    ours = {k: v for k, v in agg.items() if k.startswith("void ew_") or "reduce_" in k and "at::" not in k}
    ours_total = sum(t for _, t in ours.values())
    out = [f"# {RND}: every kernel launch of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (ncu, cold cache, serialised)",
           "", "| kernel | launches | total us | avg us | share of all | share of our kernels |", "|---|---:|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        so = f"{100 * t / ours_total:.1f}%" if k in ours else ""
        out.append(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f}% | {so} |")
    out.append("")
    out.append("Kernels from `at::` are torch's input generation and the post-run sanity checks, outside the timed region.")
    open(os.path.join(ROOT, "profiles", f"{RND}_launches_summary.md"), "w").write("\n".join(out) + "\n")
    shutil.copy(src, os.path.join(ROOT, "profiles", f"{RND}_launches_bench_steps2.csv"))

    full = json.load(open(os.path.join(ROOT, "profiles", f"{RND}_ncu_full_summary.json")))

    def gb(s):
        v, u = s.split()[0], s.split()[1]
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        return float(v) * mult

    traffic = {}
    try:  # captures of earlier rounds stay unless this round re-captured the kernel
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    prev_source = traffic.pop("source", None)
    names = {f"{RND}_wide_cfg2.ncu-rep": "ew_tile_wide_kernel_cfg2_bytes", f"{RND}_tile_cfg2.ncu-rep": "ew_tile_kernel_cfg2_bytes", f"{RND}_rows_cfg1.ncu-rep": "ew_rows_kernel_cfg1_bytes",
             f"{RND}_redrows_cfg3.ncu-rep": "reduce_rows_kernel_cfg3_bytes", f"{RND}_cols_cfg3.ncu-rep": "reduce_cols_kernel_cfg3_bytes"}
    for rep, key in names.items():
        if rep in full:
            traffic[key] = int(round(gb(full[rep]["dram__bytes_read.sum"]) + gb(full[rep]["dram__bytes_write.sum"])))
    import datetime
    fresh = [k for rep, k in names.items() if rep in full]
    traffic["source"] = (f"profiles/{RND}_ncu_full_summary.json ({datetime.date.today().isoformat()}; re-captured: {', '.join(fresh) or 'none'}): "
                         "dram__bytes_read.sum + dram__bytes_write.sum per launch of one `ncu --set full --clock-control none` "
                         "capture per kernel; algorithmic bytes: cfg2 8,589,934,592, cfg1 1,073,807,360, cfg3 2,147,614,720 (rows) / "
                         "2,147,614,720 (cols)" + (f" | earlier: {prev_source[:60]}..." if prev_source and RND not in prev_source else ""))
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))
    print("\n".join(out[:12]))


if __name__ == "__main__":
    main()
