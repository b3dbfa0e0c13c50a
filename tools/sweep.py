#!/usr/bin/env python3
"""Device-resident throughput sweep: pairs/s and per-kernel us/pair for (pairs, wave, profile on/off),
with NVML clock / power sampling at 20 ms.  usage: sweep.py "pairs,wave,profile" ...   (GPU box only)"""
import os, sys, threading, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "old-audiosync_b200")):
    sys.path.insert(0, p)
import torch
import audiosync_cuda as ac
import pynvml
pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)

class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True); self.on = True; self.s = []
    def run(self):
        while self.on:
            self.s.append((pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM),
                           pynvml.nvmlDeviceGetPowerUsage(H) / 1000.0,
                           pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_MEM)))
            time.sleep(0.02)

L = int(os.environ.get("SWEEP_L", "1440000"))
steps = int(os.environ.get("SWEEP_STEPS", "4"))
dev = torch.device("cuda", 0)
ctx = ac.Context([0])
maxp = max(int(a.split(",")[0]) for a in sys.argv[1:])
d_src = torch.empty(maxp * 2 * L, dtype=torch.float32, device=dev)
d_smp = torch.empty(maxp * L, dtype=torch.float32, device=dev)
d_res = torch.zeros(maxp * ac.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(dev)   # non-default: a NULL handle would select the library's own stream
ctx.synth_pairs(0, 0x5EED, 0, maxp, L, ac.F32, d_src.data_ptr(), d_smp.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
for a in sys.argv[1:]:
    n, wave, prof = (int(x) for x in a.split(","))
    ctx.set_wave_pairs(wave)
    def step():
        ctx.xcorr_batch_device(0, d_src.data_ptr(), d_smp.data_ptr(), n, L, ac.F32, d_res.data_ptr(), st.cuda_stream)
    for _ in range(2): step()
    torch.cuda.synchronize()
    ctx.profile_enable(bool(prof)); ctx.profile_reset()
    sm = Sampler(); sm.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps): step()
    e1.record(st); torch.cuda.synchronize()
    sm.on = False; sm.join()
    ms = e0.elapsed_time(e1)
    pr = ctx.profile_read() if prof else {}
    ctx.profile_enable(False)
    clk = sorted(x[0] for x in sm.s); pw = sorted(x[1] for x in sm.s)
    per = {k: round(1e3 * v[1] / (n * steps), 2) for k, v in pr.items() if v[1] > 0}
    print(json.dumps({"pairs": n, "wave": wave, "profile": prof, "pairs_per_s": round(n * steps / ms * 1e3, 1),
                      "us_per_pair": round(1e3 * ms / (n * steps), 2), "kernel_us_per_pair": per,
                      "sm_mhz_min_med_max": [clk[0], clk[len(clk) // 2], clk[-1]] if clk else None,
                      "power_w_med_max": [pw[len(pw) // 2], pw[-1]] if pw else None, "samples": len(clk)}), flush=True)
