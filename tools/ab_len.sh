#!/bin/bash
# A/B of variant libraries over several lengths: tools/ab_len.sh "v1 v2" "L1 L2 ..."
for L in $2; do for n in $1; do
  P=$(python -c "print(max(16, min(8192, int(512*1440000/$L))))")
  echo -n "$n L=$L: "; AUDIOSYNC_CUDA_LIB=old-audiosync_b200/variants/$n.so SWEEP_L=$L SWEEP_STEPS=3 python tools/sweep.py $P,0,1 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['us_per_pair'], d['kernel_us_per_pair'])"
done; done | tee -a gpurun_out/ab_len.log
