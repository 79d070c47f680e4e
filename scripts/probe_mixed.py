#!/usr/bin/env python
"""Mixed-dtype + on 2^27 elements: fused promotion (rc_ew_mixed.cu) vs the cast-then-op path (RC_EW_FUSED_PROMOTE=0).
GB/s counts the algorithmic bytes of the fused form: sizeof(TA) + sizeof(TB) + sizeof(K) per element."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rstsr_b200 as rt

torch.cuda.set_device(0)
dev = rt.DeviceCuda(0, rt.ROW_MAJOR, stream=torch.cuda.current_stream().cuda_stream)
tag = "two_pass" if os.environ.get("RC_EW_FUSED_PROMOTE") == "0" else "fused"
n = 1 << 27
rows = []
for (tta, nta), (ttb, ntb) in (((torch.float32, np.float32), (torch.float64, np.float64)), ((torch.float64, np.float64), (torch.int32, np.int32)),
                               ((torch.int64, np.int64), (torch.float64, np.float64)), ((torch.int32, np.int32), (torch.int64, np.int64))):
    a = (torch.rand(n, device="cuda") * 1000).to(tta)
    b = (torch.rand(n, device="cuda") * 1000).to(ttb)
    K = np.promote_types(nta, ntb)
    c = torch.empty(n, dtype=getattr(torch, K.name), device="cuda")
    ra, rb, rc = dev.wrap(a.data_ptr(), n, nta), dev.wrap(b.data_ptr(), n, ntb), dev.wrap(c.data_ptr(), n, K)
    l = rt.Layout((n,), (1,))
    for _ in range(3):
        dev.op_mutc_refa_refb("add", rc, l, ra, l, rb, l)
    torch.cuda.synchronize()
    ok = bool(torch.equal(c, a.to(c.dtype) + b.to(c.dtype)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dev.op_mutc_refa_refb("add", rc, l, ra, l, rb, l)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    nb = n * (np.dtype(nta).itemsize + np.dtype(ntb).itemsize + K.itemsize)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        torch.add(a, b, out=c)
    t1.record(); torch.cuda.synchronize()
    row = {"mode": tag, "pair": f"{np.dtype(nta).name}+{np.dtype(ntb).name}", "us": round(us, 1), "gbs": round(nb / us / 1e3), "exact": ok,
           "torch_us": round(t0.elapsed_time(t1) * 100, 1)}
    print(json.dumps(row), flush=True)
    rows.append(row)
    assert ok
json.dump(rows, open(f"gpurun_out/probe_mixed_{tag}.json", "w"), indent=1)
