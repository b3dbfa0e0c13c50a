"""NumPy (pocketfft, fp64) restatement of the hot path -- TEST INFRASTRUCTURE.

An independent second opinion on oracle/xcorr_oracle.c: same algorithm,
different FFT implementation (numpy's pocketfft instead of our FFTW stand-in).
Follows reference src/cross_correlation.c:141-142 (sizes), :164-166 (zero pad),
:232-233 (conj product), :237 (UNNORMALISED c2r => multiply numpy's irfft by
N), :52-67 (argmax semantics), :256-271 (fold + windows), :74-116 (Pearson),
:276 (NaN gate).  The alignment in the reference's dev/audiosync_sketch.py
(np.roll) differs from the C windows and is deliberately not used.

Also holds the vectorised NumPy form of the seeded synthetic generator
(SURVEY.md section 8d); tests check it bit for bit against the C generator.
"""
from __future__ import annotations

import numpy as np

MIN_CONFIDENCE = 0.95          # include/audiosync/audiosync.h:24
SAMPLE_RATE = 48000            # include/audiosync/audiosync.h:14
INTERV_SAMPLE = [s * SAMPLE_RATE for s in (3, 6, 10, 15, 20, 30)]  # src/audiosync.c:50-57


def max_abs_index(r: np.ndarray) -> int:
    """src/cross_correlation.c:52-67 (signed r[0] seed, strict >, first wins)."""
    if r.shape[0] == 1:
        return 0
    mag = np.abs(r[1:])
    mag = np.where(np.isnan(mag), -np.inf, mag)
    j = int(np.argmax(mag))            # first occurrence of the maximum
    return j + 1 if mag[j] > r[0] else 0


def pearson(x: np.ndarray, y: np.ndarray) -> float:
    """src/cross_correlation.c:74-116 (two-pass, mean-centred)."""
    n = x.shape[0]
    with np.errstate(all="ignore"):
        if n == 0:
            return float("nan")
        dx = x - np.sum(x) / n
        dy = y - np.sum(y) / n
        return float(np.sum(dx * dy) / np.sqrt(np.sum(dx * dx) * np.sum(dy * dy)))


def correlation(source: np.ndarray, sample: np.ndarray) -> np.ndarray:
    """results[] of src/cross_correlation.c:237-239 (length 2L, scaled by N)."""
    L = sample.shape[0]
    N = 2 * L
    padded = np.zeros(N, np.float64)
    padded[:L] = sample
    fa = np.fft.rfft(np.asarray(source[:N], np.float64))
    fb = np.fft.rfft(padded)
    return np.fft.irfft(fa * np.conj(fb), n=N) * N


def windows(source: np.ndarray, sample: np.ndarray, idx: int):
    """Fold + aligned windows of src/cross_correlation.c:256-271: (lag, source window, sample window)."""
    L = sample.shape[0]
    if idx >= L:                                  # :256-263
        lag = (idx % L) - L
        return lag, source[0:L + lag], sample[-lag:L]
    return idx, source[idx:idx + L], sample[0:L]  # :264-270


def peak_quality(source: np.ndarray, sample: np.ndarray, raw_index: int, peak: float, second: float):
    """Peak-quality outputs of the batched records (not in the reference, which
    only thresholds the Pearson coefficient): margin of the peak over the second peak and the
    peak normalised by the energies of the aligned windows of :256-271,
        margin = (|peak| - second) / |peak|            (0 when peak == 0)
        ncc    = peak / (N * sqrt(sum wx^2 * sum wy^2)),  N = 2L  (peak carries FFTW's factor N)."""
    source = np.asarray(source, np.float64); sample = np.asarray(sample, np.float64)
    L = sample.shape[0]
    _, wx, wy = windows(source, sample, raw_index)
    ap = abs(peak)
    margin = (ap - second) / ap if ap > 0 else (0.0 if ap == 0 else float("nan"))
    with np.errstate(all="ignore"):
        ncc = float(np.float64(peak) / (np.float64(2 * L) * np.sqrt(np.sum(wx * wx) * np.sum(wy * wy))))
    return dict(margin=float(margin), ncc=ncc)


def cross_correlation(source: np.ndarray, sample: np.ndarray):
    """Returns dict(ret, lag, coef, raw_index, peak, second, margin, ncc)."""
    source = np.asarray(source, np.float64)
    sample = np.asarray(sample, np.float64)
    L = sample.shape[0]
    r = correlation(source, sample)
    idx = max_abs_index(r)
    mag = np.abs(r).copy()
    mag[idx] = -1.0
    second = float(np.nanmax(mag)) if mag.shape[0] > 1 else 0.0
    if idx >= L:                                  # :256-263
        lag = (idx % L) - L
        wx = source[0:L + lag]
        wy = sample[-lag:L]
    else:                                         # :264-270
        lag = idx
        wx = source[lag:lag + L]
        wy = sample[0:L]
    coef = pearson(wx, wy)
    ret = -1 if coef != coef else 0               # :276
    out = dict(ret=ret, lag=int(lag), coef=coef, raw_index=idx, peak=float(r[idx]), second=second)
    out.update(peak_quality(source, sample, idx, out["peak"], second))
    return out


# ------------------------------------------------------------ synthetic pairs

_M64 = (1 << 64) - 1
_K_PAIR = 0xD1342543DE82EF95
_K_STREAM = 0xA0761D6478BD642F


def _splitmix64_scalar(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _key(seed: int, pair_id: int, stream: int) -> int:
    return (seed ^ ((pair_id * _K_PAIR) & _M64) ^ ((stream * _K_STREAM) & _M64)) & _M64


def _q(key: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(key) + idx.astype(np.uint64))
    return (h >> np.uint64(40)).astype(np.int64) - (1 << 23)


def synth_true_lag(seed: int, pair_id: int, L: int) -> int:
    return _splitmix64_scalar(_key(seed, pair_id, 2)) % (L + 1) - L // 2


def synth_pair_int(seed: int, pair_id: int, L: int):
    """Integer samples (value = int * 2**-23): (source[2L], sample[L]) as int64."""
    k0, k1 = _key(seed, pair_id, 0), _key(seed, pair_id, 1)
    tl = synth_true_lag(seed, pair_id, L)
    amp = 768 if pair_id % 4 == 3 else 102
    half = L // 2
    source = _q(k0, np.arange(half, half + 2 * L, dtype=np.int64))
    base = _q(k0, np.arange(half + tl, half + tl + L, dtype=np.int64))
    noise = _q(k1, np.arange(L, dtype=np.int64))
    sample = base + ((noise * amp) >> 10)
    return source, sample


def synth_pair(seed: int, pair_id: int, L: int, dtype=np.float64):
    s, p = synth_pair_int(seed, pair_id, L)
    scale = dtype(1.0) / dtype(8388608.0)
    return s.astype(dtype) * scale, p.astype(dtype) * scale
