"""Host-thread safety of the C ABI: handles are independent (own stream, own workspace); one handle may be shared by
several threads (launches serialise on its stream; the split-reduction workspace is guarded)."""
import threading

import numpy as np
import pytest

import rstsr_b200 as rt

from helpers import seed_of

pytestmark = pytest.mark.gpu


def work(dev, seed, out, errors):
    try:
        rng = np.random.default_rng(seed)
        for it in range(12):
            n = int(rng.integers(1000, 400000))
            a, b = rng.standard_normal(n), rng.standard_normal(n)
            ta, tb = rt.asarray(a, dev), rt.asarray(b, dev)
            c = (ta + tb) * 2.0
            if not np.array_equal(c.to_numpy(), (a + b) * 2.0):
                raise AssertionError(f"elementwise mismatch (seed {seed}, it {it})")
            s = c.sum_all()                      # split reduction: first pass + second pass through the workspace
            if abs(s - ((a + b) * 2.0).sum()) > 1e-12 * np.abs((a + b) * 2.0).sum():
                raise AssertionError(f"sum mismatch (seed {seed}, it {it})")
            m = rt.asarray(a[: (n // 64) * 64], dev).reshape([n // 64, 64])
            if not np.allclose(m.sum_axes(0).to_numpy(), a[: (n // 64) * 64].reshape(-1, 64).sum(0), rtol=1e-11, atol=1e-9):
                raise AssertionError("column sum mismatch")
            if rt.vecdot(ta, tb).to_numpy().shape != ():
                raise AssertionError("vecdot shape")
            if not rt.allclose(ta, ta):
                raise AssertionError("allclose")
            try:                                 # errors are thread-local: message must be this thread's
                ta.index_select(0, [n])
            except rt.RstsrCudaError as e:
                if e.kind != "IndexError":
                    raise AssertionError(f"wrong error kind {e.kind}")
        out.append(seed)
    except Exception as e:  # noqa: BLE001
        errors.append(repr(e))


@pytest.mark.parametrize("shared", [False, True])
def test_concurrent_host_threads(dev, shared):
    devs = [dev] * 6 if shared else [rt.DeviceCuda(0, rt.ROW_MAJOR) for _ in range(6)]
    out, errors = [], []
    threads = [threading.Thread(target=work, args=(d, seed_of("thr", shared, i), out, errors)) for i, d in enumerate(devs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if not shared:
        for d in devs:
            d.close()
    assert not errors, errors
    assert len(out) == 6
