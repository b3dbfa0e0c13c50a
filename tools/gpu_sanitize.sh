#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over a subset of the GPU parity suite that touches every kernel family:
# static four-step (shortest + headline length), runtime-radix (smooth, embedded, fp64, static rows), single-CTA, direct,
# Pearson / peak quality, session pool (async appends), mixed streams, concurrent drop-in callers, host narrowing
# (exact fp32 images through the wide Pearson instantiation, both-ways feeding, the give-up path).
mkdir -p gpurun_out
K='golden_batch_device_f32 and (144000 or 1440000) or kat_cross or second_peak_all_paths or (generic_plan_any_length and (4099 or 24000 or 250000 or 192000 or 10007)) or generic_plan_edges or (precise_mode and 144000) or peak_quality or pool_async or streams_may_be_mixed or concurrent_callers or (lossless_host_narrowing_is_bit_identical and (6000 or 1000)) or narrowing_gives_up'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/san_$tool.log
done
