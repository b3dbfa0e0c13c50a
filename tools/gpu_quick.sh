#!/bin/bash
# Quick GPU visit: selected tests + device-resident sweep with the wave pipeline on and off.
# usage: tools/gpu_quick.sh TAG [pytest -k expression]
TAG=${1:-q}; KEXPR=${2:-pipeline}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
for pl in ${PIPES:-0}; do
  echo "== AUDIOSYNC_CUDA_PIPELINE=$pl"
  AUDIOSYNC_CUDA_PIPELINE=$pl timeout 600 python tools/sweep.py ${SWEEP_ARGS:-2048,0,0 2048,0,1 2048,32,0 2048,128,0} 2>&1 | tee -a gpurun_out/sweep_$TAG.log
done
