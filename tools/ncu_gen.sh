#!/bin/bash
# ncu --set full of the three transform kernels of one wave at sample_len L for a variant library.
# usage: tools/ncu_gen.sh VARIANT L [pairs]
V=$1; L=$2; PAIRS=${3:-64}
mkdir -p gpurun_out
AUDIOSYNC_CUDA_LIB=old-audiosync_b200/variants/$V.so timeout 600 ncu --set full --clock-control none --import-source on \
  -k "regex:kernel_entry" -s 3 -c 3 -f -o gpurun_out/gen_${V}_$L python bench.py --sample-len $L --pairs $PAIRS --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency \
  > gpurun_out/gen_${V}_$L.log 2>&1; echo "$V $L rc=$?"
