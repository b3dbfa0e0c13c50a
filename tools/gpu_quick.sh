#!/bin/bash
# Quick GPU visit: selected tests + a device-resident sweep.
# usage: tools/gpu_quick.sh TAG [pytest -k expression] ; SWEEP_ARGS="pairs,wave,profile ..." overrides the sweep
TAG=${1:-q}; KEXPR=${2:-golden}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
timeout 600 python tools/sweep.py ${SWEEP_ARGS:-2048,0,0 2048,0,1 2048,32,0 2048,128,0} 2>&1 | tee -a gpurun_out/sweep_$TAG.log
