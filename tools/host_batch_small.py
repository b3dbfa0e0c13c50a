#!/usr/bin/env python3
"""Host-fed F64 batch throughput at a SHORT sample_len (many small pairs): narrowing off / lossless, pinned / pageable.
usage: host_batch_small.py [L] [pairs]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "old-audiosync_b200")):
    sys.path.insert(0, p)
import numpy as np
import audiosync_cuda as ac
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
rng = np.random.default_rng(1)
src = rng.integers(-2**20, 2**20, n * 2 * L).astype(np.float64) / 2**20
smp = rng.integers(-2**20, 2**20, n * L).astype(np.float64) / 2**20
out = {"L": L, "pairs": n, "copy_threads": ac.copy_threads()}
with ac.Context([0]) as c, ac.RealBuffer(n * 2 * L) as sb, ac.RealBuffer(n * L) as mb:
    sb.array[:] = src; mb.array[:] = smp
    ref = None
    for mem, (ps, pm) in (("pinned", (sb.ptr, mb.ptr)), ("pageable", (src.ctypes.data, smp.ctypes.data))):
        for mode, name in ((ac.NARROW_OFF, "off"), (ac.NARROW_LOSSLESS, "lossless")):
            c.set_host_narrowing(mode)
            r = c.xcorr_batch_records(ps, pm, n, L, ac.F64, ac.HOST)
            c.host_feed_stats(reset=True)
            t = time.perf_counter()
            for _ in range(3):
                r = c.xcorr_batch_records(ps, pm, n, L, ac.F64, ac.HOST)
            dt = (time.perf_counter() - t) / 3
            fed = c.host_feed_stats(reset=True)
            if ref is None: ref = r
            assert all(np.array_equal(r[f], ref[f], equal_nan=True) for f in ac.RESULT_DTYPE.names)
            out["%s_%s" % (mem, name)] = {"kpairs_per_s": round(n / dt / 1e3, 1), "host_gbs": round(n * 3 * L * 8 / dt / 1e9, 1), "fed": fed}
print(json.dumps(out))
