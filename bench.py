#!/usr/bin/env python
"""bench.py -- headline measurement of the DeviceCuda hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A STEP is one pass of the hot path over one batch of synthetic tensors: the three kernels the metric names,
on the configurations of BASELINE.json (per GPU; weak scaling, every rank owns the same-sized shard):
    cfg1   c = a + b            f64 (8192,8192) + (8192,)                     1,073,807,360 B
    cfg2   to_contig(RowMajor)  f64 (1024,1024,512) viewed transpose(2,0,1)   8,589,934,592 B   <- dominant kernel
    cfg3   sum over axis -1 and over axis 0 of f64 (16384,16384)             2 x 2,147,614,720 B
At N > 1 the rows of cfg3 are sharded, so the axis-0 sum reduces the sharded axis: rc_reduce_axes_sharded combines
the (16384,) partials inside the timed region (the one exchange step of this path) and the combined result is
checked against a torch all_reduce of the per-rank column sums.
`value` = algorithmic bytes of all ranks / max-over-ranks device time, inputs resident in HBM.
`per_config` = the other named configurations, measured in the same process with the same rules: cfg2 ColMajor,
          cfg3 f32 / max, and the SHARDED configs strong-scaled over the N ranks: cfg4 (64,64,512,512) split on output
          axis 0 and cfg5 2^33 f64 sum_all / max_all split evenly (collective and host scalar inside the timing);
          every output is verified against torch on the same data.
`e2e`   = same metric through the C ABI with HOST buffers: pinned host -> device copies of every input and
          device -> host copies of every result inside the timed region; next to it the bare copy ceiling of the
          same bytes (plain cudaMemcpyAsync both ways at once from the same buffers, all ranks concurrently).
`--impl reference`: the reference's CPU path (C/OpenMP port in oracle/, every host core) on the FULL step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "achieved HBM GB/s (and % of 8 TB/s) for broadcast add, transpose-copy, axis sum"
UNIT = "GB/s"
N1 = 8192                       # cfg1
SHP2 = (1024, 1024, 512)        # cfg2
N3 = 16384                      # cfg3
BYTES_CFG1 = 2 * N1 * N1 * 8 + N1 * 8
BYTES_CFG2 = 2 * SHP2[0] * SHP2[1] * SHP2[2] * 8
BYTES_CFG3 = N3 * N3 * 8 + N3 * 8
BYTES_STEP = BYTES_CFG1 + BYTES_CFG2 + 2 * BYTES_CFG3
WORKLOAD = ("f64: cfg1 add (8192,8192)+(8192,) | cfg2 to_contig(RowMajor) of (1024,1024,512).transpose(2,0,1) | "
            "cfg3 sum axis -1 and axis 0 of (16384,16384); per GPU")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def published_baseline():
    return None  # the reference publishes no number for this metric


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  NVML is queried
    in-process every 10 ms (a looping `nvidia-smi -lms` child was seen to stall the stream for ~2 ms per sample, i.e.
    4 % of a 40 ms timed region); nvidia-smi is only the fallback when NVML cannot be initialised."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []   # (sm_mhz, max_mhz, reasons bitmask)
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((float(mhz), float(self.max_mhz), int(mask)))
            except Exception:
                pass
            self._stop.wait(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            if self._thread is not None:
                self._thread.join(timeout=1.0)
            n = self.nvml
            names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                     ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                     ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
            reasons = set()
            for name, new, old in names:
                bit = getattr(n, new, None) or getattr(n, old, 0)
                if any(m & bit for _, _, m in self.samples):
                    reasons.add(name)
            sm = sorted(s for s, _, _ in self.samples)
            try:
                n.nvmlShutdown()
            except Exception:
                pass
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.max_mhz), "reasons": sorted(reasons),
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the C + OpenMP port of the DeviceFaer loops (oracle/), FULL step
# ------------------------------------------------------------------------------------------------------------
def host_cores():
    """CPUs this process may run on (the rayon global pool of DeviceFaer::default() sizes itself the same way)."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except Exception:
        cpus = list(range(os.cpu_count() or 1))
    return cpus


def _cpu_ranges(cpus):
    out, i = [], 0
    while i < len(cpus):
        j = i
        while j + 1 < len(cpus) and cpus[j + 1] == cpus[j] + 1:
            j += 1
        out.append(f"{cpus[i]}-{cpus[j]}" if j > i else f"{cpus[i]}")
        i = j + 1
    return ",".join(out)


class CpuPort:
    """Times the reference's CPU algorithm (rayon regimes restated with OpenMP, oracle/rstsr_oracle.c) on the FULL step:
    cfg1 (8192,8192)+(8192,); cfg2 (1024,1024,512) permuted copy; cfg3 (16384,16384) both axes -- the same shapes,
    dtypes and byte counts as the GPU arm (13.96 GB of algorithmic traffic, about a second per step on a 2-socket box).
    Thread count = every CPU in the process's affinity mask, set EXPLICITLY: torchrun exports OMP_NUM_THREADS=1.
    Conservative for the reference: the OpenMP port uses a static `omp for` where rayon spawns nested per-element
    tasks for runs >= 4096 (cpu_rayon/op_with_func.rs:57-67), so the real DeviceFaer is not faster than this."""
    SAMPLE = "full step: cfg1 (8192,8192)+(8192,); cfg2 (1024,1024,512) permuted copy; cfg3 (16384,16384) sum axis -1 and 0"

    def __init__(self):
        import ctypes
        import numpy as np
        import oracle
        from oracle import layout as OL
        self.np, self.oracle, self.OL = np, oracle, OL
        self.lib = oracle.load(native=True)  # -march=native build on the box that runs it
        self.lib.orc_num_threads.restype = ctypes.c_int
        self.cpus = host_cores()
        try:
            self.lib.orc_set_num_threads.argtypes = [ctypes.c_int]
            self.lib.orc_set_num_threads(len(self.cpus))
        except AttributeError:
            pass
        self.cores = int(self.lib.orc_num_threads())
        rng = np.random.default_rng(42)
        self.a1 = rng.random(N1 * N1)
        self.b1 = rng.random(N1)
        self.c1 = np.empty(N1 * N1)
        self.src2 = rng.random(SHP2[0] * SHP2[1] * SHP2[2])
        self.dst2 = np.empty_like(self.src2)
        self.m3 = rng.random(N3 * N3)
        self.bytes = BYTES_STEP
        # first touch of every output page outside the timed region
        self.c1[:] = 0
        self.dst2[:] = 0

    def step(self):
        import ctypes
        np, oracle, OL = self.np, self.oracle, self.OL
        cl = oracle._cl
        # cfg1: translate_to_col_major(K) + with_contig -> run 8192, outer [8192]
        la = OL.c_contig_layout([N1, N1])
        la_b, lb_b = OL.broadcast_layout(la, OL.c_contig_layout([N1]), "row")
        full = OL.translate_to_col_major([la, la_b, lb_b], "K")
        outer, run = OL.translate_to_col_major_with_contig(full)
        use = outer if run >= 16 else full
        self.lib.orc_add_par_f64(self.c1.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cl(use[0])),
                                 self.a1.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cl(use[1])),
                                 self.b1.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cl(use[2])), ctypes.c_int64(run))
        # cfg2: assign_arbitary, row-major device, strided source -> per-element two-odometer copy in parallel
        lsrc = OL.c_contig_layout(SHP2).transpose([2, 0, 1])
        ldst = OL.c_contig_layout(lsrc.shape)
        oracle.assign_arbitary(self.dst2, ldst, self.src2, lsrc, "row", parallel=True, lib=self.lib)
        # cfg3: reduce_axes regimes (a) and (b)
        l3 = OL.c_contig_layout([N3, N3])
        oracle.reduce_axes("sum", self.m3, l3, [-1], device="rayon", lib=self.lib)
        oracle.reduce_axes("sum", self.m3, l3, [0], device="rayon", lib=self.lib)

    def describe(self, value):
        return {"value": round(value, 2), "unit": UNIT, "cores": self.cores, "kind": "port", "sample": self.SAMPLE,
                "affinity": _cpu_ranges(self.cpus), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                "note": "conservative: static omp-for where rayon nests per-element tasks (op_with_func.rs:57-67)"}


def run_reference(args, rank):
    if rank != 0:
        return
    port = CpuPort()
    for _ in range(max(args.warmup, 1)):
        port.step()
    # the whole run must end within a few minutes: at most `steps` steps and at most ~100 s of timed work
    t0 = time.perf_counter()
    done = 0
    while done < args.steps and (done < 2 or time.perf_counter() - t0 < 100.0):
        port.step()
        done += 1
    dt = (time.perf_counter() - t0) / done
    v = port.bytes / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": done, "warmup": max(args.warmup, 1), "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": port.SAMPLE, "bytes_per_step": BYTES_STEP},
            "cpu_baseline": port.describe(v),
            "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "C/OpenMP port of the DeviceFaer loops (the Rust reference cannot be built: no rustc in the image); "
                    "one host, all its cores, whatever --gpus says"}
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------------------
def run_product(args, rank, local_rank, world):
    import numpy as np
    import torch

    import rstsr_b200 as rt
    from rstsr_b200 import Layout

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream().cuda_stream
    dev = rt.DeviceCuda(local_rank, rt.ROW_MAJOR, stream=stream)

    comm = None
    peer_window = False
    if world > 1:
        uid = [rt.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = rt.Comm(dev, world, rank, uid[0])
        peer_window = comm.info()[2]

    def wrap(t, dt=np.float64):
        return dev.wrap(t.data_ptr(), t.numel(), dt)

    g = torch.Generator(device="cuda")
    g.manual_seed(42 + rank)
    f64 = dict(dtype=torch.float64, device="cuda")
    a1 = torch.rand(N1 * N1, generator=g, **f64)
    b1 = torch.rand(N1, generator=g, **f64)
    c1 = torch.empty(N1 * N1, **f64)
    src2 = torch.rand(SHP2[0] * SHP2[1] * SHP2[2], generator=g, **f64)
    dst2 = torch.empty_like(src2)
    m3 = torch.rand(N3 * N3, generator=g, **f64)
    o3r = torch.empty(N3, **f64)
    o3c = torch.empty(N3, **f64)
    ra1, rb1, rc1, rs2, rd2, rm3, ror, roc = map(wrap, (a1, b1, c1, src2, dst2, m3, o3r, o3c))

    la1 = Layout((N1, N1), (N1, 1))
    lb1 = Layout((N1, N1), (0, 1))
    lsrc2 = Layout((SHP2[2], SHP2[0], SHP2[1]), (1, SHP2[1] * SHP2[2], SHP2[2]))
    ldst2 = Layout.contig(lsrc2.shape, rt.ROW_MAJOR)
    lm3 = Layout((N3, N3), (N3, 1))
    lo3 = Layout((N3,), (1,))

    ev = {k: [] for k in ("cfg1", "cfg2", "cfg3_rows", "cfg3_cols")}

    def timed(name, fn, record):
        if record:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            ev[name].append((e0, e1))
        else:
            fn()

    def axis0_sum():
        if comm is None:
            dev.reduce_axes_into("sum", rm3, lm3, [0], roc, lo3)
        else:  # rows are sharded across ranks: local partial + cross-GPU combine in one entry point
            comm.reduce_axes_sharded("sum", rm3, lm3, [0], N3 * world, roc, lo3)

    def step(record=False):
        timed("cfg1", lambda: dev.op_mutc_refa_refb("add", rc1, la1, ra1, la1, rb1, lb1), record)
        timed("cfg2", lambda: dev.assign_arbitary(rd2, ldst2, rs2, lsrc2), record)
        timed("cfg3_rows", lambda: dev.reduce_axes_into("sum", rm3, lm3, [-1], ror, lo3), record)
        timed("cfg3_cols", axis0_sum, record)

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.25)
    launches0 = dev.launch_count()
    sync_all()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step()                      # the headline region carries no per-op events (8 event records cost ~16 us a step)
    t1.record()
    sync_all()
    launches = dev.launch_count() - launches0
    for _ in range(args.steps):     # attribution pass: the same K steps again, each op bracketed by events
        step(record=True)
    sync_all()
    elapsed = torch.tensor([t0.elapsed_time(t1) * 1e-3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed.item())
    clk = clocks.stop() if clocks else None

    # ---- the timed results are checked (outside the timed region) on EVERY rank ----
    assert torch.equal(c1.view(N1, N1), a1.view(N1, N1) + b1)
    probe = src2.view(*SHP2)[5:7].permute(2, 0, 1).contiguous()
    assert torch.equal(dst2.view(SHP2[2], SHP2[0], SHP2[1])[:, 5:7, :], probe)
    ref_r = m3.view(N3, N3).sum(dim=1)
    assert float(((o3r - ref_r).abs() / ref_r.abs()).max()) < 1e-12
    ref_c = m3.view(N3, N3).sum(dim=0)
    if dist is not None:  # the combined column sums: every rank must hold the all-reduced result
        dist.all_reduce(ref_c)
    assert float(((o3c - ref_c).abs() / ref_c.abs()).max()) < 1e-12, "axis-0 sum (combined over ranks) mismatch"
    if dist is not None:  # ... and bitwise the same one on every rank (rank-ordered fold)
        same = o3c.clone()
        dist.broadcast(same, src=0)
        if peer_window:
            assert torch.equal(same, o3c), "combined result differs between ranks"
    checks = {"cfg1": "bit-exact vs torch", "cfg2": "bit-exact vs torch (slab)", "cfg3_rows": "rel 1e-12 vs torch",
              "cfg3_cols": ("rel 1e-12 vs torch all_reduce of the per-rank sums; bitwise equal on all ranks"
                            if world > 1 else "rel 1e-12 vs torch")}

    peak, peak_kind = measured_peak()
    per_op = {"_note": "separate attribution pass of the same K steps with CUDA events around every op (max over ranks); "
                       "an op's time includes the write-back of its predecessor's dirty L2 lines"}
    for k, pairs in ev.items():
        ms = sum(e0.elapsed_time(e1) for e0, e1 in pairs) / max(len(pairs), 1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        ms_min = ms
        if dist is not None:
            tmin = t.clone()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            ms_min = float(tmin.item())
        ms = float(t.item())
        nbytes = {"cfg1": BYTES_CFG1, "cfg2": BYTES_CFG2, "cfg3_rows": BYTES_CFG3, "cfg3_cols": BYTES_CFG3}[k]
        gbs = nbytes / (ms * 1e-3) / 1e9
        per_op[k] = {"us": round(ms * 1e3, 1), "gbs": round(gbs, 1), "frac_measured": round(gbs / peak, 4),
                     "frac_8TBs": round(gbs / 8000, 4), "check": checks[k]}
        if dist is not None:  # fastest rank: the spread is GPU-to-GPU variation (no collective in cfg1 / cfg2 / cfg3_rows)
            per_op[k]["us_fastest_rank"] = round(ms_min * 1e3, 1)
    del a1, b1, c1, src2, dst2, m3, o3r, o3c, probe, ref_r, ref_c

    # ---- the other named configurations, same process, same rules ----
    per_config = None
    if not args.no_per_config:
        per_config = run_per_config(args, rank, world, dev, comm, rt, np, torch, dist, peak)
    torch.cuda.empty_cache()

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region ----
    e2e = run_e2e(args, dev, rt, np, torch, dist, comm, world)
    if dist is not None:
        if comm is not None:
            comm.close()
        dist.barrier()
        dist.destroy_process_group()

    if rank != 0:
        return
    value = world * BYTES_STEP * args.steps / elapsed / 1e9
    dom = per_op["cfg2"]
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get("ew_tile_wide_kernel_cfg2_bytes", tj.get("ew_tile_kernel_cfg2_bytes"))
        traffic_src = tj.get("source")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(elapsed / args.steps * 1e3, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": published_baseline(), "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (every array >= 1 GiB vs 126 MB L2)",
                   "collective": ("rc_reduce_axes_sharded: " + ("one-kernel NVLink peer-window combine" if peer_window else
                                  "ncclAllReduce") + " of the (16384,) axis-0 partial") if world > 1 else "none",
                   "bytes_per_step_per_gpu": BYTES_STEP},
        "pct_of_8TBs": round(value / world / 8000 * 100, 1),
        "pct_of_measured_peak": round(value / world / peak * 100, 1),
        "per_op": per_op,
        "per_config": per_config,
        "roofline": {"bound": "hbm", "kernel": "ew_tile_wide_kernel<u64, 128, 32, 256> (cfg2 permuted copy)",
                     "achieved": dom["gbs"], "peak": peak, "peak_kind": peak_kind + " (burst copy, MEASURED_PEAKS.json)",
                     "unit": "GB/s", "frac": round(dom["gbs"] / peak, 4), "traffic": traffic,
                     "traffic_source": traffic_src, "algorithmic_bytes": BYTES_CFG2},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            port = CpuPort()
            port.step()
            t = time.perf_counter()
            reps = 0
            while reps < 2 or (time.perf_counter() - t < 10 and reps < 20):
                port.step()
                reps += 1
            dt = (time.perf_counter() - t) / reps
            line["cpu_baseline"] = port.describe(port.bytes / dt / 1e9)
        except Exception as exc:  # the checker is test infrastructure: its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {exc}"}
    emit(line)


def run_per_config(args, rank, world, dev, comm, rt, np, torch, dist, peak):
    """cfg2 ColMajor and the six remaining cfg3 cases (weak: every rank the full named shape), cfg4 and cfg5
    STRONG-scaled over the ranks (the named total split evenly).  CUDA events on the launching stream, >= 3 warm-ups,
    max over ranks, working sets far above L2; every output verified against torch on the same data."""
    from rstsr_b200 import Layout
    rows = {}
    iters = max(3, min(args.steps, 10))

    def wrap(t, dt=np.float64):
        return dev.wrap(t.data_ptr(), t.numel(), dt)

    def timeit(fn, n=iters, warmup=3):
        for _ in range(warmup):
            fn()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e-3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def report(name, total_bytes, sec, scaling, check):
        gbs = total_bytes / sec / 1e9
        rows[name] = {"scaling": scaling, "us": round(sec * 1e6, 1), "gbs_total": round(gbs, 1),
                      "gbs_per_gpu": round(gbs / world, 1), "frac_measured": round(gbs / world / peak, 4),
                      "frac_8TBs": round(gbs / world / 8000, 4), "bytes_total": int(total_bytes), "check": check}

    g = torch.Generator(device="cuda")
    g.manual_seed(4242 + rank)
    f64 = dict(dtype=torch.float64, device="cuda")

    # ---- cfg2, ColMajor target (inner axis contiguous on both sides: vectorised copy, no transpose tile) ----
    N = SHP2[0] * SHP2[1] * SHP2[2]
    src = torch.rand(N, generator=g, **f64)
    dst = torch.empty(N, **f64)
    lsrc = Layout((SHP2[2], SHP2[0], SHP2[1]), (1, SHP2[1] * SHP2[2], SHP2[2]))
    ldst = Layout.contig(lsrc.shape, rt.COL_MAJOR)
    rs, rd = wrap(src), wrap(dst)
    sec = timeit(lambda: dev.assign_arbitary(rd, ldst, rs, lsrc))
    # F-contiguous (512,1024,1024) == C-contiguous (1024,1024,512) of the axes reversed
    want = src.view(*SHP2)[3:5].permute(2, 0, 1)          # [k, i, j] for i in 3:5
    got = dst.view(SHP2[1], SHP2[0], SHP2[2])[:, 3:5, :]  # memory [j][i][k]
    assert torch.equal(got.permute(2, 1, 0), want), "cfg2 ColMajor mismatch"
    report("cfg2_colmajor f64 to_contig(ColMajor) of (1024,1024,512).transpose(2,0,1)", world * 2 * N * 8, sec, "weak",
           "bit-exact vs torch (slab)")
    del src, dst, want, got

    # ---- cfg3: f32 sum / max and f64 max, both axes (f64 sum is the headline step) ----
    for dt, npdt, ops in ((torch.float32, np.float32, ("sum", "max")), (torch.float64, np.float64, ("max",))):
        m = torch.rand(N3 * N3, generator=g, dtype=dt, device="cuda")
        o = torch.empty(N3, dtype=dt, device="cuda")
        rm, ro = wrap(m, npdt), wrap(o, npdt)
        lm, lo = Layout((N3, N3), (N3, 1)), Layout((N3,), (1,))
        es = m.element_size()
        for op in ops:
            for axis in (0, -1):
                sec = timeit(lambda: dev.reduce_axes_into(op, rm, lm, [axis], ro, lo))
                ref = getattr(m.view(N3, N3), op)(dim=axis)
                ref = ref.values if op == "max" else ref
                if op == "max":
                    assert torch.equal(o, ref), f"cfg3 {dt} max axis {axis} mismatch"
                    chk = "bit-exact vs torch"
                else:  # f32 sums: tolerance 1e-5 relative (north star), both sides accumulate in f32
                    assert float(((o - ref).abs() / ref.abs()).max()) < 1e-5, f"cfg3 f32 sum axis {axis} mismatch"
                    chk = "rel 1e-5 vs torch"
                report(f"cfg3 {str(dt)[6:]} {op} axis {axis} (16384,16384)", world * (N3 * N3 * es + N3 * es), sec, "weak", chk)
        del m, o, ref

    # ---- cfg4: strong scaling, sharded on output axis 0: rank r owns a_r, c_r = (64/g,64,512,512) C-contiguous and
    #      b_r = its local C-contiguous (64, 64/g, 512, 512) block viewed permuted (1,0,3,2) ----
    D0, D1, D2, D3 = 64, 64, 512, 512
    if D0 % world == 0:
        ni = D0 // world
        loc = ni * D1 * D2 * D3
        a = torch.rand(loc, generator=g, **f64)
        bl = torch.rand(loc, generator=g, **f64)
        c = torch.empty(loc, **f64)
        ra, rb, rc = wrap(a), wrap(bl), wrap(c)
        lc = Layout.contig((ni, D1, D2, D3), rt.ROW_MAJOR)
        lbp = Layout((ni, D1, D2, D3), (D2 * D3, ni * D2 * D3, 1, D3))
        sec = timeit(lambda: dev.op_mutc_refa_refb("add", rc, lc, ra, lc, rb, lbp), n=max(3, iters // 2))
        bview = bl.view(D1, ni, D2, D3).permute(1, 0, 3, 2)
        for j in (0, D1 // 2 + 1, D1 - 1):
            assert torch.equal(c.view(ni, D1, D2, D3)[:, j], a.view(ni, D1, D2, D3)[:, j] + bview[:, j]), "cfg4 mismatch"
        report("cfg4 f64 c = a + b.transpose(1,0,3,2) (64,64,512,512), axis 0 sharded", 3 * D0 * D1 * D2 * D3 * 8, sec,
               "strong", "bit-exact vs torch (3 slabs per rank)")
        v = torch.rand(D2 * D3, generator=g, **f64)
        rv = wrap(v)
        lv = Layout((ni, D1, D2, D3), (0, 0, D3, 1))
        sec = timeit(lambda: dev.op_mutc_refa_refb("mul", rc, lc, ra, lc, rv, lv), n=max(3, iters // 2))
        assert torch.equal(c.view(ni, D1, D2 * D3)[ni - 1], a.view(ni, D1, D2 * D3)[ni - 1] * v), "cfg4b mismatch"
        report("cfg4b f64 c = a * v, v broadcast (0,0,512,1), axis 0 sharded", 2 * D0 * D1 * D2 * D3 * 8 + world * D2 * D3 * 8,
               sec, "strong", "bit-exact vs torch (slab)")
        del a, bl, c, v, bview

    # ---- cfg5: 2^33 f64 (64 GiB) split evenly; per-GPU partial + cross-GPU combine + host scalar, all timed ----
    total = 1 << 33
    free_b, _ = torch.cuda.mem_get_info()
    per = total // world
    if per * 8 + (8 << 30) <= free_b:
        x = torch.rand(per, generator=g, **f64)
        # a planted maximum on one rank makes the max check meaningful
        if rank == world - 1:
            x[per // 3] = 1.5
        rx = wrap(x)
        lx = Layout((per,), (1,))
        for op in ("sum", "max"):
            if comm is None:
                fn = lambda: dev.reduce_all(op, rx, lx)
            else:
                fn = lambda: comm.reduce_all_sharded(op, rx, lx, total)
            sec = timeit(fn, n=max(3, iters // 2))
            got = fn()
            if op == "sum":
                ref = torch.stack([x.view(1 << 10, -1).sum(dim=1).sum()])
                if dist is not None:
                    dist.all_reduce(ref)
                assert abs(got - float(ref.item())) <= 1e-12 * abs(float(ref.item())), "cfg5 sum_all mismatch"
                chk = "rel 1e-12 vs torch" + (" all_reduce" if world > 1 else "")
            else:
                ref = torch.stack([x.max()])
                if dist is not None:
                    dist.all_reduce(ref, op=dist.ReduceOp.MAX)
                assert got == float(ref.item()) == 1.5, "cfg5 max_all mismatch"
                chk = "bit-exact vs torch" + (" all_reduce(MAX)" if world > 1 else "")
            report(f"cfg5 f64 {op}_all of 2^33 elements (64 GiB) split over the ranks, incl. combine + host scalar",
                   total * 8, sec, "strong", chk)
        del x
    else:
        rows["cfg5"] = {"skipped": f"{per * 8 >> 30} GiB per rank does not fit the free {free_b >> 30} GiB"}

    # ---- the exchange step alone: ranks in lock-step, 200 back-to-back all-reduces of the cfg3 partial (128 KiB) ----
    if comm is not None:
        buf = torch.rand(N3, generator=g, **f64)
        rb = wrap(buf)
        had_peer = comm.info()[2]
        for label, enable in (("peer window", True), ("ncclAllReduce", False)):
            if enable and not had_peer:
                continue
            comm.set_peer_window(enable)   # same call on every rank, between the same two collectives
            sec = timeit(lambda: comm.all_reduce("max", rb), n=200, warmup=10)
            rows[f"exchange only: all-reduce of 16384 f64, {label}"] = {"us": round(sec * 1e6, 2), "scaling": "latency",
                                                                        "check": "see tests/test_gpu_multi.py"}
        comm.set_peer_window(had_peer)
    return rows


def pinned_alloc(rt, torch, numel, node):
    """float64 torch view of a pinned staging buffer placed on NUMA node `node` (rc_host_alloc_on_node)."""
    import ctypes
    ptr, bound = rt.DeviceCuda.host_alloc(numel * 8, node)
    buf = (ctypes.c_double * numel).from_address(ptr)
    t = torch.frombuffer(buf, dtype=torch.float64, count=numel)
    return t, ptr, bound


def run_e2e(args, dev, rt, np, torch, dist, comm, world):
    """Same step through the C ABI with HOST data: every input is copied from pinned host memory, every result
    is copied back, all inside the timed region.  Three handles of the same GPU (upload / compute / download
    streams, ordered with rc_device_wait) pipeline the step in chunks along each config's outermost axis, so
    PCIe uploads, kernels and PCIe downloads overlap (full-duplex link): the step costs ~max(H2D, D2H) instead
    of their sum.  Timed with CUDA events on the stream `dev` is bound to; the pipeline is fenced against it.
    The staging buffers live on the GPU's own NUMA node (rc_host_alloc_on_node).  `ceiling` = the same bytes moved
    by bare cudaMemcpyAsync both ways at once, all ranks concurrently: what the host links allow at this N."""
    from rstsr_b200 import Layout
    steps = max(1, min(args.steps, 3))
    up = rt.DeviceCuda(dev.ordinal, rt.ROW_MAJOR)
    cp = rt.DeviceCuda(dev.ordinal, rt.ROW_MAJOR)
    dn = rt.DeviceCuda(dev.ordinal, rt.ROW_MAJOR)
    ccomm = None
    if comm is not None:  # the collective runs on the compute handle's stream
        uid = [rt.Comm.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ccomm = rt.Comm(cp, world, dist.get_rank(), uid[0])
    node = dev.numa_node() if not args.no_numa else -1
    n2 = SHP2[0] * SHP2[1] * SHP2[2]
    ptrs, bounds = [], []

    def pin(numel, fill):
        t, ptr, bound = pinned_alloc(rt, torch, numel, node)
        ptrs.append(ptr)
        bounds.append(bound)
        if fill:
            t.copy_(torch.rand(numel, dtype=torch.float64))
        return t

    h_a1 = pin(N1 * N1, True); h_b1 = pin(N1, True); h_c1 = pin(N1 * N1, False)
    h_s2 = pin(n2, True); h_d2 = pin(n2, False)
    h_m3 = pin(N3 * N3, True); h_or = pin(N3, False); h_oc = pin(N3, False)
    d_a1 = cp.uninit_impl(np.float64, N1 * N1); d_b1 = cp.uninit_impl(np.float64, N1)
    d_c1 = cp.uninit_impl(np.float64, N1 * N1)
    d_s2 = cp.uninit_impl(np.float64, n2); d_d2 = cp.uninit_impl(np.float64, n2)
    d_m3 = cp.uninit_impl(np.float64, N3 * N3); d_or = cp.uninit_impl(np.float64, N3); d_oc = cp.uninit_impl(np.float64, N3)
    cp.synchronize()
    lm3 = Layout((N3, N3), (N3, 1)); lo3 = Layout((N3,), (1,))
    CH1, CH2, CH3 = 4, 16, 4

    def step():
        up.wait(cp); cp.wait(dn)  # do not overwrite inputs / outputs the previous step still uses
        # cfg1: row chunks of c = a + b
        up.h2d_async(d_b1, h_b1.data_ptr(), N1 * 8)
        r = N1 // CH1
        for k in range(CH1):
            off = k * r * N1
            up.h2d_async(d_a1, h_a1.data_ptr() + off * 8, r * N1 * 8, off * 8)
            cp.wait(up)
            cp.op_mutc_refa_refb("add", d_c1, Layout((r, N1), (N1, 1), off), d_a1, Layout((r, N1), (N1, 1), off),
                                 d_b1, Layout((r, N1), (0, 1)))
            dn.wait(cp)
            dn.d2h_async(h_c1.data_ptr() + off * 8, d_c1, r * N1 * 8, off * 8)
        # cfg2: chunks along the source's outermost axis i; each produces the slab dst[:, i0:i1, :]
        ci = SHP2[0] // CH2
        slab = SHP2[1] * SHP2[2]
        for k in range(CH2):
            i0 = k * ci
            up.h2d_async(d_s2, h_s2.data_ptr() + i0 * slab * 8, ci * slab * 8, i0 * slab * 8)
            cp.wait(up)
            lsrc = Layout((SHP2[2], ci, SHP2[1]), (1, slab, SHP2[2]), i0 * slab)
            ldst = Layout((SHP2[2], ci, SHP2[1]), (SHP2[0] * SHP2[1], SHP2[1], 1), i0 * SHP2[1])
            cp.assign_arbitary(d_d2, ldst, d_s2, lsrc)
            dn.wait(cp)
            pitch = SHP2[0] * SHP2[1] * 8
            dn.d2h_2d_async(h_d2.data_ptr() + i0 * SHP2[1] * 8, pitch, d_d2, i0 * SHP2[1] * 8, pitch,
                            ci * SHP2[1] * 8, SHP2[2])
        # cfg3: upload in row blocks, reduce once the matrix is resident
        r = N3 // CH3
        for k in range(CH3):
            up.h2d_async(d_m3, h_m3.data_ptr() + k * r * N3 * 8, r * N3 * 8, k * r * N3 * 8)
        cp.wait(up)
        cp.reduce_axes_into("sum", d_m3, lm3, [-1], d_or, lo3)
        if ccomm is None:
            cp.reduce_axes_into("sum", d_m3, lm3, [0], d_oc, lo3)
        else:
            ccomm.reduce_axes_sharded("sum", d_m3, lm3, [0], N3 * world, d_oc, lo3)
        dn.wait(cp)
        dn.d2h_async(h_or.data_ptr(), d_or, N3 * 8)
        dn.d2h_async(h_oc.data_ptr(), d_oc, N3 * 8)

    def fence_start():
        for h in (up, cp, dn):
            h.wait(dev)

    def fence_end():
        for h in (up, cp, dn):
            dev.wait(h)

    def timed_region(body, reps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fence_start()
        for _ in range(reps):
            body()
        fence_end()
        e1.record()
        torch.cuda.synchronize()
        sec = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return float(sec.item())

    fence_start(); step(); fence_end()
    sec = timed_region(step, steps)
    # the host results are the real thing
    assert torch.equal(h_c1.view(N1, N1), h_a1.view(N1, N1) + h_b1)
    want = h_s2.view(*SHP2)[700:702].permute(2, 0, 1).contiguous()
    assert torch.equal(h_d2.view(SHP2[2], SHP2[0], SHP2[1])[:, 700:702, :], want)
    ref_r = h_m3.view(N3, N3).sum(dim=1)
    assert float(((h_or - ref_r).abs() / ref_r.abs()).max()) < 1e-12
    ref_c = h_m3.view(N3, N3).sum(dim=0)
    if dist is not None:  # gloo is not initialised: combine the reference through the GPU
        ref_c_d = ref_c.cuda()
        dist.all_reduce(ref_c_d)
        ref_c = ref_c_d.cpu()
    assert float(((h_oc - ref_c).abs() / ref_c.abs()).max()) < 1e-12, "e2e axis-0 sum (combined over ranks) mismatch"
    h2d_bytes = (h_a1.numel() + h_b1.numel() + h_s2.numel() + h_m3.numel()) * 8
    d2h_bytes = (h_c1.numel() + h_d2.numel() + h_or.numel() + h_oc.numel()) * 8

    # ---- bare ceiling of the host links: the same bytes, plain copies, both directions at once ----
    def bare():
        up.wait(dn)
        up.h2d_async(d_a1, h_a1.data_ptr(), N1 * N1 * 8)
        up.h2d_async(d_s2, h_s2.data_ptr(), n2 * 8)
        up.h2d_async(d_m3, h_m3.data_ptr(), N3 * N3 * 8)
        dn.d2h_async(h_c1.data_ptr(), d_c1, N1 * N1 * 8)
        dn.d2h_async(h_d2.data_ptr(), d_d2, n2 * 8)

    fence_start(); bare(); fence_end()
    sec_bare = timed_region(bare, 2) / 2
    if ccomm is not None:
        ccomm.close()
    del d_a1, d_b1, d_c1, d_s2, d_d2, d_m3, d_or, d_oc
    cp.synchronize()
    del h_a1, h_b1, h_c1, h_s2, h_d2, h_m3, h_or, h_oc, want, ref_r, ref_c
    for ptr in ptrs:
        rt.DeviceCuda.host_free(ptr)
    ms = sec / steps * 1e3
    return {"value": round(world * BYTES_STEP * steps / sec / 1e9, 2), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
            "d2h_bytes_per_step": d2h_bytes, "steps": steps, "ms_per_step": round(ms, 2),
            "ceiling": {"ms_per_step": round(sec_bare * 1e3, 2),
                        "h2d_gbs_per_gpu": round(h2d_bytes / sec_bare / 1e9, 1),
                        "d2h_gbs_per_gpu": round(d2h_bytes / sec_bare / 1e9, 1),
                        "what": "bare cudaMemcpyAsync of the step's bytes, H2D and D2H concurrently, all ranks at once"},
            "frac_of_ceiling": round(sec_bare * 1e3 / ms, 3),
            "numa": {"gpu_node": node, "placement": {0: "none", 1: "mbind", 2: "first-touch"}.get(max(bounds) if bounds else 0)},
            "check": "host outputs: cfg1 / cfg2 bit-exact, cfg3 rel 1e-12 (axis 0 vs the all-reduced reference at N>1)",
            "path": "pinned host -> rc_memcpy_h2d_async -> rc_op_mutc_refa_refb / rc_assign_arbitary / "
                    "rc_reduce_axes_into / rc_reduce_axes_sharded -> rc_memcpy(2d)_d2h_async -> pinned host; "
                    "upload/compute/download handles ordered by rc_device_wait, chunked 4/16/4"}


_SAVED_STDOUT = None


def _guard_stdout():
    global _SAVED_STDOUT
    if _SAVED_STDOUT is None:
        sys.stdout.flush()
        _SAVED_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    """print the one JSON line on the real stdout"""
    sys.stdout.flush()
    if _SAVED_STDOUT is not None:
        os.dup2(_SAVED_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip the per_config block (cfg2 ColMajor, cfg3 f32/max, cfg4, cfg5)")
    ap.add_argument("--no-numa", action="store_true", help="e2e staging buffers without NUMA placement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl != "reference" and world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it (one process per GPU); the children print the line
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    # stdout carries exactly one JSON line: whatever a library writes to fd 1 meanwhile (NCCL prints its version
    # banner there when the box sets NCCL_DEBUG=VERSION) is sent to stderr; emit() restores fd 1 for the line.
    _guard_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_product(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
