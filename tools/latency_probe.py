#!/usr/bin/env python3
"""Single-pair device-resident latency (BASELINE config 3), p50/p95 over 300 calls rotating through 16 resident pairs.
usage: latency_probe.py   (GPU box only; knobs through the environment, e.g. AUDIOSYNC_CUDA_NO_PDL=1)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "old-audiosync_b200")):
    sys.path.insert(0, p)
import torch
import audiosync_cuda as ac
L = int(os.environ.get("SWEEP_L", "1440000")); n = 16; reps = 300
dev = torch.device("cuda", 0)
ctx = ac.Context([0])
d_src = torch.empty(n * 2 * L, dtype=torch.float32, device=dev)
d_smp = torch.empty(n * L, dtype=torch.float32, device=dev)
d_res = torch.zeros(n * ac.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(dev)
ctx.synth_pairs(0, 0x5EED, 0, n, L, ac.F32, d_src.data_ptr(), d_smp.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for i in range(20 + reps):
    k = i % n
    if i >= 20: ev[i - 20][0].record(st)
    ctx.xcorr_batch_device(0, d_src.data_ptr() + k * 2 * L * 4, d_smp.data_ptr() + k * L * 4, 1, L, ac.F32,
                           d_res.data_ptr(), st.cuda_stream)
    if i >= 20: ev[i - 20][1].record(st)
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
res = d_res.cpu().numpy().view(ac.RESULT_DTYPE)
print(json.dumps({"L": L, "p50_us": round(t[len(t) // 2], 2), "p95_us": round(t[int(len(t) * 0.95)], 2), "min_us": round(t[0], 2),
                  "lag": int(res["lag"][0]), "pdl": os.environ.get("AUDIOSYNC_CUDA_NO_PDL") is None}))
