"""The reference's UNMODIFIED orchestrator (src/audiosync.c: audiosync_run, the interval loop
of :226-259) driven by fake capture/download threads (oracle/harness/fake_io.c, the reader
protocol of src/ffmpeg_pipe.c), linked

  * against the reference's own src/cross_correlation.c + FFT stand-in  -> audiosync_harness_cpu
  * against libaudiosync_cuda.so, nothing else changed                  -> audiosync_harness_gpu

The second binary has exactly three undefined symbols that resolve into the GPU library:
cross_correlation, fftw_alloc_real, fftw_free -- the drop-in boundary of SURVEY 8b.  Both are
built by oracle/Makefile where /root/reference exists and travel to the GPU box prebuilt.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import capi  # noqa: E402

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
CPU_BIN = os.path.join(REF_DIR, "audiosync_harness_cpu")
GPU_BIN = os.path.join(REF_DIR, "audiosync_harness_gpu")
SEED, L = 0x5EED, 1440000


def _write_pair(tmp_path, pair_id):
    src, smp = capi.synth_pair(SEED, pair_id, L)
    ps, pm = tmp_path / ("src%d.f64" % pair_id), tmp_path / ("smp%d.f64" % pair_id)
    src.astype("<f8").tofile(ps); smp.astype("<f8").tofile(pm)
    return src, smp, str(ps), str(pm)


def _run(binary, ps, pm, debug=0):
    r = subprocess.run([binary, ps, pm, str(debug)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"ret=(-?\d+) lag=(-?\d+)", r.stdout)
    assert m, r.stdout
    return int(m.group(1)), int(m.group(2)), r.stderr


@pytest.mark.skipif(not os.path.exists(CPU_BIN), reason="oracle/_ref harness not built (needs /root/reference)")
@pytest.mark.parametrize("pair_id", [0, 3])
def test_cpu_harness_agrees_with_restated_loop(tmp_path, pair_id):
    """Pins oracle_interval_loop (and audiosync_cuda.interval_loop, which mirrors it) on the
    reference's real audiosync_run: same return value and same reported lag."""
    src, smp, ps, pm = _write_pair(tmp_path, pair_id)
    ret, lag, _ = _run(CPU_BIN, ps, pm)
    want = capi.interval_loop(src, smp)
    assert (ret, lag) == (want["final_ret"], want["final_lag"])
    if pair_id % 4 != 3:          # clean pair: accepted at the first interval, lag in ms
        assert ret == 0 and lag == round(capi.synth_true_lag(SEED, pair_id, L) * 1000.0 / 48000.0)
    else:                         # noisy pair: every interval is rejected, lag stays in frames
        assert ret == -1 and lag == capi.synth_true_lag(SEED, pair_id, L)


@pytest.mark.gpu
@pytest.mark.parametrize("pair_id", [0, 3, 5])
def test_unmodified_audiosync_run_on_the_gpu_library(tmp_path, pair_id):
    if not (os.path.exists(CPU_BIN) and os.path.exists(GPU_BIN)):
        pytest.skip("oracle/_ref harness binaries not present")
    _, _, ps, pm = _write_pair(tmp_path, pair_id)
    cpu = _run(CPU_BIN, ps, pm, debug=1)
    gpu = _run(GPU_BIN, ps, pm, debug=1)
    assert gpu[:2] == cpu[:2]
    # the LOG line of src/cross_correlation.c:278 appears once per evaluated interval in both
    pat = re.compile(r"(-?\d+) frames of delay with a confidence of (-?[\d.]+)")
    lc, lg = pat.findall(cpu[2]), pat.findall(gpu[2])
    assert len(lc) == len(lg) >= 1
    for (fc, cc), (fg, cg) in zip(lc, lg):
        assert fc == fg and abs(float(cc) - float(cg)) <= 1e-4
