#!/bin/bash
# A/B on one GPU box: alternates the variant libraries, REPS times each, per-kernel us/pair.
# usage: tools/ab.sh "nameA nameB ..." [REPS] [sweep args]
NAMES=$1; REPS=${2:-3}; ARGS=${3:-1024,0,1}
mkdir -p gpurun_out
for r in $(seq $REPS); do for n in $NAMES; do
  echo -n "$n: "; AUDIOSYNC_CUDA_LIB=old-audiosync_b200/variants/$n.so python tools/sweep.py $ARGS 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['us_per_pair'], d['kernel_us_per_pair'], d['sm_mhz_min_med_max'][1])"
done; done | tee -a gpurun_out/ab.log
