"""Deterministic HARD input pairs for the parity suite -- TEST INFRASTRUCTURE.

The seeded generator of SURVEY 8(d) only makes white-noise pairs whose peak stands >= 0.97 above
everything else.  These builders make the inputs an fp32 transform can get wrong: two candidate
lags whose correlation values differ by 3e-4 .. 1e-2 of the peak ("echo": the sample occurs
twice in the source, the second time slightly attenuated; "periodic": a repeating pattern plus a
little noise, so every multiple of the period is nearly as good), the extreme lags 0, +-1, L-1,
the fold boundary idx == L and idx == L + 1 (reference src/cross_correlation.c:256-271) and the
all-zero sample (:276).  Everything is integer arithmetic on splitmix64 streams (values are
int * 2**-23 with |int| < 2**24: exact in fp32 AND fp64), so the same arrays are rebuilt bit for
bit on the GPU box from the few parameters stored in tests/golden/hard.json.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import xcorr_numpy as xn  # noqa: E402

SCALE = 1.0 / 8388608.0


def _stream(seed: int, stream: int, n: int, shift: int) -> np.ndarray:
    """n integers in [-2**(23-shift), 2**(23-shift))."""
    return xn._q(xn._key(seed, 0xC0FFEE, stream), np.arange(n, dtype=np.int64)) >> shift


def _place(dst: np.ndarray, src: np.ndarray, d: int) -> None:
    """dst[d + n] += src[n] for every n with 0 <= d + n < len(dst)."""
    n0 = max(0, -d)
    n1 = min(src.shape[0], dst.shape[0] - d)
    if n1 > n0:
        dst[d + n0:d + n1] += src[n0:n1]


def _finish(source: np.ndarray, sample: np.ndarray, dtype):
    assert np.abs(source).max() < (1 << 24) and np.abs(sample).max() < (1 << 24)
    return (source.astype(dtype) * dtype(SCALE)), (sample.astype(dtype) * dtype(SCALE))


def echo_pair_int(seed: int, L: int, d1: int, d2: int, k4096: int):
    """The sample occurs in the source at lag d1 (full amplitude) and at lag d2 (amplitude
    k4096 / 4096), over a quiet independent background."""
    s = _stream(seed, 0, L, 1)                       # |s| < 2**22
    source = _stream(seed, 1, 2 * L, 3)              # background, |.| < 2**20
    _place(source, s, d1)
    _place(source, (s * k4096) >> 12, d2)
    return source, s


def periodic_pair_int(seed: int, L: int, d1: int, period: int, kn: int):
    """A pattern of `period` frames repeated through source and sample, plus noise of relative
    amplitude kn / 2**20 that matches only at lag d1."""
    pat = _stream(seed, 2, period, 2)                # |pat| < 2**21
    noise = (_stream(seed, 3, L, 1) * kn) >> 20
    idx = np.arange(2 * L, dtype=np.int64)
    source = pat[(idx - d1) % period] + ((_stream(seed, 4, 2 * L, 1) * kn) >> 20)
    sample = pat[np.arange(L, dtype=np.int64) % period] + noise
    # the sample's own noise sits in the source at lag d1 (replacing the independent noise there)
    lo, hi = max(0, d1), min(2 * L, d1 + L)
    source[lo:hi] = pat[(np.arange(lo, hi) - d1) % period] + noise[lo - d1:hi - d1]
    return source, sample


def lag_pair_int(seed: int, L: int, lag: int):
    """sample = the source at `lag` plus 5 % noise (positive or negative lag, any magnitude < L)."""
    source = _stream(seed, 5, 2 * L, 1)
    sample = (_stream(seed, 6, L, 1) * 51) >> 10
    n = np.arange(L, dtype=np.int64)
    ok = (n + lag >= 0) & (n + lag < 2 * L)
    sample[ok] += source[n[ok] + lag]
    return source, sample


def impulse_pair_int(L: int, i_src: int, i_smp: int):
    """One impulse each: r has a single non-zero entry at (i_src - i_smp) mod 2L."""
    source = np.zeros(2 * L, np.int64); sample = np.zeros(L, np.int64)
    source[i_src] = 1 << 22
    sample[i_smp] = 1 << 22
    return source, sample


def zero_sample_pair_int(seed: int, L: int):
    return _stream(seed, 7, 2 * L, 1), np.zeros(L, np.int64)


def build(case: dict, dtype=np.float64):
    """(source, sample) of a hard.json case, as `dtype` arrays."""
    k, L, p = case["kind"], case["L"], case["params"]
    if k == "echo":
        s, m = echo_pair_int(p["seed"], L, p["d1"], p["d2"], p["k4096"])
    elif k == "periodic":
        s, m = periodic_pair_int(p["seed"], L, p["d1"], p["period"], p["kn"])
    elif k == "lag":
        s, m = lag_pair_int(p["seed"], L, p["lag"])
    elif k == "impulse":
        s, m = impulse_pair_int(L, p["i_src"], p["i_smp"])
    elif k == "zero_sample":
        s, m = zero_sample_pair_int(p["seed"], L)
    else:
        raise ValueError(k)
    return _finish(s, m, dtype)


def r_at(source_int: np.ndarray, sample_int: np.ndarray, j: int) -> float:
    """N * sum_n source[(n + j) mod N] * sample[n] (the reference's r[j]) from the integers, in fp64."""
    L = sample_int.shape[0]
    N = 2 * L
    idx = (np.arange(L, dtype=np.int64) + j) % N
    return float(N) * float(np.dot(source_int[idx].astype(np.float64) * SCALE, sample_int.astype(np.float64) * SCALE))
