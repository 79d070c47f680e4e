#!/usr/bin/env python
"""bench.py -- headline measurement of the DeviceCuda hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A STEP is one pass of the hot path over one batch of synthetic tensors: the three kernels the metric names,
on the configurations of BASELINE.json (per GPU; weak scaling, every rank owns the same-sized shard):
    cfg1   c = a + b            f64 (8192,8192) + (8192,)                     1,073,807,360 B
    cfg2   to_contig(RowMajor)  f64 (1024,1024,512) viewed transpose(2,0,1)   8,589,934,592 B   <- dominant kernel
    cfg3   sum over axis -1 and over axis 0 of f64 (16384,16384)             2 x 2,147,614,720 B
At N > 1 the rows of cfg3 are sharded, so the axis-0 sum reduces the sharded axis: its (16384,) partial output
is combined with an NCCL all-reduce inside the timed region (the one exchange step of this path).
`value` = algorithmic bytes of all ranks / max-over-ranks device time, inputs resident in HBM.
`e2e`   = same metric through the C ABI with HOST buffers: pinned host -> device copies of every input and
          device -> host copies of every result inside the timed region.
`--impl reference`: the reference's CPU path (C/OpenMP port in oracle/, all host threads) on a bounded sample.
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
# reference arm / cpu baseline: the C + OpenMP port of the DeviceFaer loops (oracle/), bounded sample
# ------------------------------------------------------------------------------------------------------------
class CpuPort:
    """Times the reference's CPU algorithm (rayon regimes restated with OpenMP, oracle/rstsr_oracle.c) on a
    bounded sample of the step: 1/8 of every configuration along its outermost axis, so the mix of kernels is
    the step's own (cfg1 (1024,8192)+(8192,); cfg2 (128,1024,512); cfg3 (2048,16384), both axes)."""
    SAMPLE = "1/8 of each config along its outermost axis: cfg1 (1024,8192)+(8192,); cfg2 (128,1024,512); cfg3 (2048,16384) both axes"

    def __init__(self):
        import numpy as np
        import oracle
        from oracle import layout as OL
        self.np, self.oracle, self.OL = np, oracle, OL
        self.lib = oracle.load(native=True)  # -march=native build on the box that runs it
        self.lib.orc_num_threads.restype = __import__("ctypes").c_int
        self.cores = int(self.lib.orc_num_threads())
        rng = np.random.default_rng(42)
        self.r1 = N1 // 8
        self.a1 = rng.random(self.r1 * N1)
        self.b1 = rng.random(N1)
        self.c1 = np.empty(self.r1 * N1)
        self.s2 = (SHP2[0] // 8, SHP2[1], SHP2[2])
        self.src2 = rng.random(self.s2[0] * self.s2[1] * self.s2[2])
        self.dst2 = np.empty_like(self.src2)
        self.r3 = N3 // 8
        self.m3 = rng.random(self.r3 * N3)
        self.bytes = (2 * self.a1.size * 8 + N1 * 8) + 2 * self.src2.size * 8 + 2 * (self.m3.size * 8 + N3 * 8)
        # first touch of every output page outside the timed region
        self.c1[:] = 0
        self.dst2[:] = 0

    def step(self):
        import ctypes
        np, oracle, OL = self.np, self.oracle, self.OL
        cl = oracle._cl
        # cfg1: translate_to_col_major(K) + with_contig -> run 8192, outer [8192]
        la = OL.c_contig_layout([self.r1, N1])
        la_b, lb_b = OL.broadcast_layout(la, OL.c_contig_layout([N1]), "row")
        full = OL.translate_to_col_major([la, la_b, lb_b], "K")
        outer, run = OL.translate_to_col_major_with_contig(full)
        use = outer if run >= 16 else full
        self.lib.orc_add_par_f64(self.c1.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cl(use[0])),
                                 self.a1.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cl(use[1])),
                                 self.b1.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cl(use[2])), ctypes.c_int64(run))
        # cfg2: assign_arbitary, row-major device, strided source -> per-element two-odometer copy in parallel
        lsrc = OL.c_contig_layout(self.s2).transpose([2, 0, 1])
        ldst = OL.c_contig_layout(lsrc.shape)
        oracle.assign_arbitary(self.dst2, ldst, self.src2, lsrc, "row", parallel=True, lib=self.lib)
        # cfg3: reduce_axes regimes (a) and (b)
        l3 = OL.c_contig_layout([self.r3, N3])
        oracle.reduce_axes("sum", self.m3, l3, [-1], device="rayon", lib=self.lib)
        oracle.reduce_axes("sum", self.m3, l3, [0], device="rayon", lib=self.lib)


def run_reference(args, rank):
    if rank != 0:
        return
    port = CpuPort()
    for _ in range(args.warmup):
        port.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.step()
    dt = (time.perf_counter() - t0) / args.steps
    v = port.bytes / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": port.SAMPLE},
            "cpu_baseline": {"value": round(v, 2), "unit": UNIT, "cores": port.cores, "kind": "port",
                             "sample": port.SAMPLE},
            "e2e": {"value": round(v, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "C/OpenMP port of the DeviceFaer loops (the Rust reference cannot be built: no rustc in the image)"}
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
    if world > 1:
        uid = [rt.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = rt.Comm(dev, world, rank, uid[0])

    def wrap(t):
        return dev.wrap(t.data_ptr(), t.numel(), np.float64)

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
        dev.reduce_axes_into("sum", rm3, lm3, [0], roc, lo3)
        if comm is not None:  # rows are sharded across ranks: combine the partial column sums
            comm.all_reduce("sum", roc, N3)

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
        step(record=True)
    t1.record()
    sync_all()
    launches = dev.launch_count() - launches0
    elapsed = torch.tensor([t0.elapsed_time(t1) * 1e-3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed.item())
    clk = clocks.stop() if clocks else None

    # sanity of the timed results (outside the timed region)
    assert torch.equal(c1.view(N1, N1), a1.view(N1, N1) + b1)
    probe = src2.view(*SHP2)[5:7].permute(2, 0, 1).contiguous()
    assert torch.equal(dst2.view(SHP2[2], SHP2[0], SHP2[1])[:, 5:7, :], probe)
    ref_r = m3.view(N3, N3).sum(dim=1)
    assert float(((o3r - ref_r).abs() / ref_r.abs()).max()) < 1e-12

    per_op = {}
    for k, pairs in ev.items():
        ms = sum(e0.elapsed_time(e1) for e0, e1 in pairs) / max(len(pairs), 1)
        nbytes = {"cfg1": BYTES_CFG1, "cfg2": BYTES_CFG2, "cfg3_rows": BYTES_CFG3, "cfg3_cols": BYTES_CFG3}[k]
        per_op[k] = {"us": round(ms * 1e3, 1), "gbs": round(nbytes / (ms * 1e-3) / 1e9, 1)}

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region ----
    e2e = run_e2e(args, dev, rt, np, torch, dist, comm, world)
    if dist is not None:
        if comm is not None:
            comm.close()
        dist.barrier()
        dist.destroy_process_group()

    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    value = world * BYTES_STEP * args.steps / elapsed / 1e9
    dom = per_op["cfg2"]
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("ew_tile_kernel_cfg2_bytes")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(elapsed / args.steps * 1e3, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": published_baseline(), "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (every array >= 1 GiB vs 126 MB L2)",
                   "collective": "NCCL all-reduce of the (16384,) axis-0 partial at N>1" if world > 1 else "none",
                   "bytes_per_step_per_gpu": BYTES_STEP},
        "pct_of_8TBs": round(value / world / 8000 * 100, 1),
        "pct_of_measured_peak": round(value / world / peak * 100, 1),
        "per_op": per_op,
        "roofline": {"bound": "hbm", "kernel": "ew_tile_kernel<FIdentity<u64>> (cfg2 permuted copy)",
                     "achieved": dom["gbs"], "peak": peak, "peak_kind": peak_kind + " (burst copy, MEASURED_PEAKS.json)",
                     "unit": "GB/s", "frac": round(dom["gbs"] / peak, 4), "traffic": traffic,
                     "algorithmic_bytes": BYTES_CFG2},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            port = CpuPort()
            port.step()
            t = time.perf_counter()
            reps = 0
            while reps < 2 or (time.perf_counter() - t < 8 and reps < 20):
                port.step()
                reps += 1
            dt = (time.perf_counter() - t) / reps
            line["cpu_baseline"] = {"value": round(port.bytes / dt / 1e9, 2), "unit": UNIT, "cores": port.cores,
                                    "kind": "port", "sample": port.SAMPLE}
        except Exception as exc:  # the checker is test infrastructure: its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {exc}"}
    emit(line)


def run_e2e(args, dev, rt, np, torch, dist, comm, world):
    """Same step through the C ABI with HOST data: every input is copied from pinned host memory, every result
    is copied back, all inside the timed region.  Three handles of the same GPU (upload / compute / download
    streams, ordered with rc_device_wait) pipeline the step in chunks along each config's outermost axis, so
    PCIe uploads, kernels and PCIe downloads overlap (full-duplex link): the step costs ~max(H2D, D2H) instead
    of their sum.  Timed with CUDA events on the stream `dev` is bound to; the pipeline is fenced against it."""
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
    pin = dict(dtype=torch.float64, pin_memory=True)
    n2 = SHP2[0] * SHP2[1] * SHP2[2]
    h_a1 = torch.rand(N1 * N1, **pin); h_b1 = torch.rand(N1, **pin); h_c1 = torch.empty(N1 * N1, **pin)
    h_s2 = torch.rand(n2, **pin); h_d2 = torch.empty(n2, **pin)
    h_m3 = torch.rand(N3 * N3, **pin); h_or = torch.empty(N3, **pin); h_oc = torch.empty(N3, **pin)
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
        cp.reduce_axes_into("sum", d_m3, lm3, [0], d_oc, lo3)
        if ccomm is not None:
            ccomm.all_reduce("sum", d_oc, N3)
        dn.wait(cp)
        dn.d2h_async(h_or.data_ptr(), d_or, N3 * 8)
        dn.d2h_async(h_oc.data_ptr(), d_oc, N3 * 8)

    def fence_start():
        for h in (up, cp, dn):
            h.wait(dev)

    def fence_end():
        for h in (up, cp, dn):
            dev.wait(h)

    fence_start(); step(); fence_end()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    fence_start()
    for _ in range(steps):
        step()
    fence_end()
    e1.record()
    torch.cuda.synchronize()
    sec = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    sec = float(sec.item())
    # the host results are the real thing
    assert torch.equal(h_c1.view(N1, N1), h_a1.view(N1, N1) + h_b1)
    want = h_s2.view(*SHP2)[700:702].permute(2, 0, 1).contiguous()
    assert torch.equal(h_d2.view(SHP2[2], SHP2[0], SHP2[1])[:, 700:702, :], want)
    ref_r = h_m3.view(N3, N3).sum(dim=1)
    assert float(((h_or - ref_r).abs() / ref_r.abs()).max()) < 1e-12
    if ccomm is None:
        ref_c = h_m3.view(N3, N3).sum(dim=0)
        assert float(((h_oc - ref_c).abs() / ref_c.abs()).max()) < 1e-12
    h2d_bytes = (h_a1.numel() + h_b1.numel() + h_s2.numel() + h_m3.numel()) * 8
    d2h_bytes = (h_c1.numel() + h_d2.numel() + h_or.numel() + h_oc.numel()) * 8
    if ccomm is not None:
        ccomm.close()
    return {"value": round(world * BYTES_STEP * steps / sec / 1e9, 2), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
            "d2h_bytes_per_step": d2h_bytes, "steps": steps, "ms_per_step": round(sec / steps * 1e3, 2),
            "path": "pinned host -> rc_memcpy_h2d_async -> rc_op_mutc_refa_refb / rc_assign_arbitary / "
                    "rc_reduce_axes_into -> rc_memcpy(2d)_d2h_async -> pinned host; upload/compute/download handles "
                    "ordered by rc_device_wait, chunked 4/16/4"}


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
