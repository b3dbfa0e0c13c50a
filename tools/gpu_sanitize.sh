mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_batch_device_f32 and (144000 or 1440000) or kat_cross or second_peak_all_paths or wave_pipeline_matches" > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$?"; tail -4 gpurun_out/san_$tool.log
done
for L in 144000 288000 480000 720000 960000; do echo "L=$L"; SWEEP_L=$L python tools/sweep.py 2048,0,1 2>&1 | tail -1; done | tee gpurun_out/sweep_lengths.log
