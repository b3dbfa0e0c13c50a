#!/bin/bash
# ncu --set full capture of one launch of the row kernel for each variant library named.
# usage: tools/ncu_kb.sh "name1 name2" [kernel regex] [pairs]
NAMES=$1; KRE=${2:-RowFused}; PAIRS=${3:-64}
mkdir -p gpurun_out
for n in $NAMES; do
  AUDIOSYNC_CUDA_LIB=old-audiosync_b200/variants/$n.so timeout 600 ncu --set full --clock-control none --import-source on \
    -k "regex:fft_kernel_entry" -s ${SKIP:-1} -c ${COUNT:-1} -f -o gpurun_out/kb_$n python bench.py --pairs $PAIRS --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency \
    > gpurun_out/kb_$n.log 2>&1; echo "$n rc=$?"
done
