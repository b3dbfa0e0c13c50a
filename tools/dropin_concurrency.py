#!/usr/bin/env python3
"""Drop-in cross_correlation() under concurrent callers: calls/s for 1, 2, 4, 8 threads, pinned and
pageable host doubles, at two interval lengths.  (GPU box only.)   usage: dropin_concurrency.py [devices]
AUDIOSYNC_CUDA_DEVICES=all spreads the callers' slots over every GPU."""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "old-audiosync_b200")):
    sys.path.insert(0, p)
import numpy as np
import audiosync_cuda as ac
from oracle import capi

out = {"devices_env": os.environ.get("AUDIOSYNC_CUDA_DEVICES", os.environ.get("AUDIOSYNC_CUDA_DEVICE", "0")),
       "slots_env": os.environ.get("AUDIOSYNC_CUDA_DROPIN_SLOTS", "3 per device (default)"), "runs": []}
for L in (144000, 1440000):
    pairs = [capi.synth_pair(0x5EED, i, L) for i in range(4)]
    truth = [capi.synth_true_lag(0x5EED, i, L) for i in range(4)]
    for memory in ("pinned", "pageable"):
        bufs = []
        for s, m in pairs:
            if memory == "pinned":
                sb, mb = ac.RealBuffer(2 * L), ac.RealBuffer(L)
                sb.array[:] = s; mb.array[:] = m
                bufs.append((sb.ptr, mb.ptr, sb, mb))
            else:
                a, b = np.ascontiguousarray(s), np.ascontiguousarray(m)
                bufs.append((a.ctypes.data, b.ctypes.data, a, b))
        base = None
        for nthreads in (1, 2, 4, 8):
            calls = max(8, int(2e8 // (3 * L * 8)) * 4)         # per thread
            bad = [0]
            def work(t):
                for k in range(calls):
                    i = (t + k) % 4
                    ret, lag, coef = ac.cross_correlation_ptr(bufs[i][0], bufs[i][1], L)
                    if ret != 0 or lag != truth[i]:
                        bad[0] += 1
            work(0) if nthreads == 1 else None                   # warm-up
            th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
            t0 = time.perf_counter()
            [t.start() for t in th]; [t.join() for t in th]
            dt = time.perf_counter() - t0
            rate = nthreads * calls / dt
            base = base or rate
            out["runs"].append({"L": L, "host_memory": memory, "threads": nthreads, "calls_per_s": round(rate, 1),
                                "vs_one_thread": round(rate / base, 2), "h2d_gbs": round(rate * 3 * L * 8 / 1e9, 1), "wrong": bad[0]})
            print(json.dumps(out["runs"][-1]), flush=True)
        for b in bufs:
            if memory == "pinned":
                b[2].free(); b[3].free()
print(json.dumps(out))
