#!/bin/bash
# Per-length device-resident throughput: us per pair and per 144,000 frames, with per-kernel event timing.
# usage: tools/sweep_lengths.sh "L1 L2 ..." [TAG]   (pairs scaled so every length moves about the same bytes)
LS=${1:-"144000 1440000"}; TAG=${2:-len}
mkdir -p gpurun_out
for L in $LS; do
  P=$(python -c "print(max(16, min(8192, int(768*1440000/$L))))")
  echo -n "L=$L pairs=$P: "
  SWEEP_L=$L SWEEP_STEPS=3 python tools/sweep.py $P,0,1 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); L=$L
print(json.dumps({'L':L,'us_per_pair':d['us_per_pair'],'us_per_144k_frames':round(d['us_per_pair']*144000/L,3),'kernels':d['kernel_us_per_pair'],'mhz':d['sm_mhz_min_med_max'][1]}))"
done | tee -a gpurun_out/sweep_lengths_$TAG.log
