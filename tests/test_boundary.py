"""The C-ABI boundary without a GPU: the library loads, exports every symbol
include/audiosync_cuda.h declares, and fails loudly (no CPU fallback) when
there is no CUDA device."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "audiosync_cuda.h")


@pytest.fixture(scope="module")
def ac():
    subprocess.run(["make", "-C", os.path.join(ROOT, "old-audiosync_b200")], check=True,
                   stdout=subprocess.DEVNULL)
    import audiosync_cuda
    return audiosync_cuda


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\(", text)
    return sorted({n for n in names if n.startswith(("audiosync_cuda_", "fftw_"))
                   or n in ("cross_correlation", "pearson_coefficient")})


def test_header_and_binding_agree(ac):
    assert declared_functions() == sorted(ac.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(ac):
    lib = ac.lib()
    for name in declared_functions():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", ac.lib_path()], check=True,
                         capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared_functions()) <= exported
    # nothing else leaks out of the shared object as a strong text symbol
    assert exported == set(declared_functions())


def test_kernel_image_is_sm100a_only(ac):
    out = subprocess.run(["cuobjdump", "-lelf", ac.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_no_oracle_or_cufft_in_product(ac):
    out = subprocess.run(["ldd", ac.lib_path()], capture_output=True, text=True).stdout
    assert "cufft" not in out and "oracle" not in out and "fftw" not in out
    strings = subprocess.run(["nm", "-D", ac.lib_path()], capture_output=True, text=True).stdout
    assert "oracle_" not in strings and "cufft" not in strings.lower()
    pkg = os.path.join(ROOT, "old-audiosync_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "cufft" not in src.lower().replace("cufft, or", ""), f


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="checks the no-device behaviour")
def test_fails_loudly_without_a_device(ac, capfd):
    src = np.zeros(20); smp = np.ones(10)
    with pytest.raises(ac.AudiosyncCudaError):
        ac.cross_correlation(src, smp)
    err = capfd.readouterr().err
    assert "audiosync:" in err and "no CPU fallback" in err
    v = ac.pearson_coefficient(np.arange(4.0), np.arange(4.0))
    assert v != v                                   # NaN, never a CPU-computed value
    with pytest.raises(ac.AudiosyncCudaError):
        ac.Context()


def test_missing_library_raises(ac, monkeypatch):
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['AUDIOSYNC_CUDA_LIB']='/nonexistent.so';"
            "import audiosync_cuda as a\n"
            "try:\n a.lib()\nexcept a.AudiosyncCudaError as e:\n print('RAISED')\n") % os.path.join(ROOT, "old-audiosync_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "RAISED" in out.stdout


def test_allocator_names_work_without_a_device(ac):
    lib = ac.lib()
    p = lib.fftw_alloc_real(1000)
    assert p and p % 64 == 0
    import ctypes
    arr = (ctypes.c_double * 1000).from_address(p)
    arr[0] = 1.5; arr[999] = -2.5
    assert arr[0] == 1.5 and arr[999] == -2.5
    lib.fftw_free(p)
    lib.fftw_free(None)


def test_frames_to_ms_matches_oracle(ac):
    from oracle import capi
    for lag in (0, 1, -1, 23, 24, 25, -24, -25, 704463, -263641, 71, 72, 73, 1439999, -1440000):
        assert ac.frames_to_ms(lag) == capi.lib().oracle_frames_to_ms(lag), lag
