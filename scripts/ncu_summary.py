"""Summarise .ncu-rep captures (read here, no GPU) into profiles/*.md|json.  Usage: python scripts/ncu_summary.py"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        d[h] = (v, u)
    return d


def main():
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    summary = {}
    for rep in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
        if not rep.endswith(".ncu-rep"):
            continue
        d = raw(os.path.join(ROOT, "gpurun_out", rep))
        name = d.get("Kernel Name", ("?", ""))[0]
        s = {"kernel": name}
        for k in WANT:
            if k in d:
                s[k] = f"{d[k][0]} {d[k][1]}".strip()
        summary[rep] = s
    # also next to the captures, so a GPU-box run can return the (small) summary and drop the (large) reports
    out_name = sys.argv[1] if len(sys.argv) > 1 else "r01_ncu_full_summary.json"  # a second name keeps later captures apart
    for d_out in (out_dir, os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(d_out, out_name), "w") as f:
            json.dump(summary, f, indent=1)
    for rep, s in summary.items():
        print(rep)
        for k, v in s.items():
            print("   ", k, "=", v)


if __name__ == "__main__":
    main()
