#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the path kernels.
# usage: tools/gpu_round.sh TAG      (afterwards, here: tools/make_profile_summary.py TAG 64 128 r02)
TAG=${1:-vX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc > gpurun_out/nproc_$TAG.txt; lscpu | head -20 >> gpurun_out/nproc_$TAG.txt
cp MEASURED_PEAKS.json gpurun_out/MEASURED_PEAKS_seen.json 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo "reference rc=$?"
# launch list: 128 pairs per launch, every kernel of the bench (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --pairs 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture: 64 pairs per launch, the four path kernels of the timed step
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:fft_kernel_entry|pearson' -s 4 -c 4 -f -o gpurun_out/prof_$TAG \
  python bench.py --pairs 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/ | tail -12
