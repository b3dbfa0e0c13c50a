"""Host-side double -> float conversion of libaudiosync_cuda (audiosync_cuda_host_narrow): the
values and the exactness verdict that decides whether an F64 host batch may cross PCIe as fp32
(include/audiosync_cuda.h, host narrowing).  No GPU needed."""
import warnings

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ac():
    import audiosync_cuda
    return audiosync_cuda


@pytest.mark.parametrize("n", [0, 1, 7, 63, 64, 65, 1000, 1 << 20, (1 << 21) + 13])
def test_exact_inputs_round_trip(ac, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32).astype(np.float64)
    out, exact = ac.host_narrow(x)
    assert exact
    assert np.array_equal(out, x.astype(np.float32))


@pytest.mark.parametrize("n", [1, 65, 1000, (1 << 20) + 5])
@pytest.mark.parametrize("bad", [np.nan, 1e-50, 1e300, -1e300, 1.0 + 2.0 ** -30, float(np.pi), 2.0 ** -150])
def test_one_inexact_value_anywhere_is_noticed(ac, n, bad):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(n).astype(np.float32).astype(np.float64)
    for pos in sorted({0, n // 3, n // 2, n - 1}):
        y = x.copy()
        y[pos] = bad
        out, exact = ac.host_narrow(y)
        assert not exact, (pos, bad)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            want = y.astype(np.float32)            # round to nearest even, like the device's conversion
        assert np.array_equal(out, want, equal_nan=True)


def test_special_values_that_are_exact(ac):
    x = np.array([0.0, -0.0, np.inf, -np.inf, 2.0 ** -149, -(2.0 ** -126), 3.4028234663852886e38, 1.0, -32768.0 / 32768.0] * 9)
    out, exact = ac.host_narrow(x)
    assert exact
    assert np.array_equal(out, x.astype(np.float32))
    assert np.array_equal(np.signbit(out), np.signbit(x))


def test_pcm16_and_pcm24_audio_is_always_exact(ac):
    rng = np.random.default_rng(3)
    s16 = rng.integers(-32768, 32768, 100_000).astype(np.float64) / 32768.0
    s24 = rng.integers(-(1 << 23), 1 << 23, 100_000).astype(np.float64) / float(1 << 23)
    assert ac.host_narrow(s16)[1] and ac.host_narrow(s24)[1]
    s32 = rng.integers(-(1 << 31), 1 << 31, 100_000).astype(np.float64) / float(1 << 31)
    assert not ac.host_narrow(s32)[1]
