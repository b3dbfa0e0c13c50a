"""ctypes access to the compiled oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two shared objects are wrapped:

* ``oracle/liboracle.so``      -- our C restatement (xcorr_oracle.c) + FFT shim.
* ``oracle/_ref/libaudiosync_ref.so`` -- the reference's UNMODIFIED
  ``src/cross_correlation.c`` compiled against the same shim (built only where
  ``/root/reference`` exists; the built file travels to the GPU box).

``kind`` in the loaders' return value says which one ran so that reports can be
labelled "port" (restatement) or "reference" (compiled reference source).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None
_REF: Optional[C.CDLL] = None


class OracleExtra(C.Structure):
    _fields_ = [("raw_index", C.c_long), ("peak", C.c_double),
                ("second", C.c_double), ("r0", C.c_double)]


def build(quiet: bool = True) -> None:
    """(Re)build liboracle.so and, where the reference checkout exists, oracle/_ref."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_cross_correlation.restype = C.c_int
        L.oracle_cross_correlation.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double),
                                               C.c_size_t, C.POINTER(C.c_long),
                                               C.POINTER(C.c_double), C.POINTER(OracleExtra)]
        L.oracle_pearson.restype = C.c_double
        L.oracle_pearson.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t]
        L.oracle_max_abs_index.restype = C.c_size_t
        L.oracle_max_abs_index.argtypes = [C.POINTER(C.c_double), C.c_size_t]
        L.oracle_accept.restype = C.c_int
        L.oracle_accept.argtypes = [C.c_int, C.c_double]
        L.oracle_frames_to_ms.restype = C.c_long
        L.oracle_frames_to_ms.argtypes = [C.c_long]
        L.oracle_interval_loop.restype = C.c_int
        L.oracle_interval_loop.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(C.c_int), C.POINTER(C.c_long),
                                           C.POINTER(C.c_double), C.POINTER(C.c_int),
                                           C.POINTER(C.c_int), C.POINTER(C.c_long)]
        L.synth_true_lag.restype = C.c_long
        L.synth_true_lag.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t]
        L.synth_pair_i32.restype = None
        L.synth_pair_i32.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.synth_pair_f64.restype = None
        L.synth_pair_f64.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t,
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.synth_pair_f32.restype = None
        L.synth_pair_f32.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t,
                                     C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_synth_batch.restype = C.c_int
        L.oracle_synth_batch.argtypes = [C.c_uint64, C.c_uint64, C.c_size_t, C.c_size_t, C.c_int,
                                         C.POINTER(C.c_long), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]
        L.oracle_backend.restype = C.c_char_p
        _LIB = L
    return _LIB


def ref_lib() -> Optional[C.CDLL]:
    """The compiled reference source, or None if oracle/_ref was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libaudiosync_ref.so")
        if not os.path.exists(path):
            if os.path.exists("/root/reference/src/cross_correlation.c"):
                build()
            if not os.path.exists(path):
                return None
        R = C.CDLL(path)
        # reference include/audiosync/cross_correlation.h:10-11,24-25
        R.cross_correlation.restype = C.c_int
        R.cross_correlation.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t,
                                        C.POINTER(C.c_long), C.POINTER(C.c_double)]
        R.pearson_coefficient.restype = C.c_double
        R.pearson_coefficient.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _REF = R
    return _REF


def backend() -> str:
    return lib().oracle_backend().decode()


# ---------------------------------------------------------------- wrappers

def cross_correlation(source: np.ndarray, sample: np.ndarray):
    """Restatement (port). Returns dict(ret, lag, coef, raw_index, peak, second, r0)."""
    source = np.ascontiguousarray(source, dtype=np.float64)
    sample = np.ascontiguousarray(sample, dtype=np.float64)
    L = sample.shape[0]
    assert source.shape[0] >= 2 * L
    lag, coef, ex = C.c_long(0), C.c_double(0.0), OracleExtra()
    ret = lib().oracle_cross_correlation(_dptr(source), _dptr(sample), L,
                                         C.byref(lag), C.byref(coef), C.byref(ex))
    out = dict(ret=ret, lag=lag.value, coef=coef.value, raw_index=ex.raw_index,
               peak=ex.peak, second=ex.second, r0=ex.r0)
    from . import xcorr_numpy
    out.update(xcorr_numpy.peak_quality(source, sample, ex.raw_index, ex.peak, ex.second))
    return out


def ref_cross_correlation(source: np.ndarray, sample: np.ndarray):
    """The reference's own compiled cross_correlation(). Returns (ret, lag, coef)."""
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref/libaudiosync_ref.so not built")
    # no copies here: the reference never writes to its inputs (src/cross_correlation.c:130-132),
    # and this call sits inside bench.py's timed loops
    source = np.ascontiguousarray(source, dtype=np.float64)
    sample = np.ascontiguousarray(sample, dtype=np.float64)
    L = sample.shape[0]
    lag, coef = C.c_long(0), C.c_double(0.0)
    ret = R.cross_correlation(_dptr(source), _dptr(sample), L, C.byref(lag), C.byref(coef))
    return ret, lag.value, coef.value


def pearson(x: np.ndarray, y: np.ndarray) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    assert x.shape == y.shape
    return lib().oracle_pearson(_dptr(x), _dptr(y), x.shape[0])


def ref_pearson(x: np.ndarray, y: np.ndarray) -> float:
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref/libaudiosync_ref.so not built")
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = x.shape[0]
    return R.pearson_coefficient(x.ctypes.data, x.ctypes.data + 8 * n,
                                 y.ctypes.data, y.ctypes.data + 8 * n)


def max_abs_index(arr: np.ndarray) -> int:
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    return int(lib().oracle_max_abs_index(_dptr(arr), arr.shape[0]))


def interval_loop(source: np.ndarray, sample: np.ndarray):
    source = np.ascontiguousarray(source, dtype=np.float64)
    sample = np.ascontiguousarray(sample, dtype=np.float64)
    assert source.shape[0] >= 2880000 and sample.shape[0] >= 1440000
    rets = (C.c_int * 6)(); lags = (C.c_long * 6)(); coefs = (C.c_double * 6)()
    succ = (C.c_int * 6)(); fr = C.c_int(0); fl = C.c_long(0)
    n = lib().oracle_interval_loop(_dptr(source), _dptr(sample), rets, lags, coefs, succ,
                                   C.byref(fr), C.byref(fl))
    return dict(n=n, rets=list(rets)[:n], lags=list(lags)[:n], coefs=list(coefs)[:n],
                succ=list(succ)[:n], final_ret=fr.value, final_lag=fl.value)


def synth_pair(seed: int, pair_id: int, L: int, dtype=np.float64):
    if dtype == np.float64:
        src = np.empty(2 * L, np.float64); smp = np.empty(L, np.float64)
        lib().synth_pair_f64(seed, pair_id, L, _dptr(src), _dptr(smp))
    elif dtype == np.float32:
        src = np.empty(2 * L, np.float32); smp = np.empty(L, np.float32)
        lib().synth_pair_f32(seed, pair_id, L, src.ctypes.data_as(C.POINTER(C.c_float)),
                             smp.ctypes.data_as(C.POINTER(C.c_float)))
    elif dtype == np.int32:
        src = np.empty(2 * L, np.int32); smp = np.empty(L, np.int32)
        lib().synth_pair_i32(seed, pair_id, L, src.ctypes.data_as(C.POINTER(C.c_int32)),
                             smp.ctypes.data_as(C.POINTER(C.c_int32)))
    else:
        raise TypeError(dtype)
    return src, smp


def synth_true_lag(seed: int, pair_id: int, L: int) -> int:
    return int(lib().synth_true_lag(seed, pair_id, L))


def synth_batch(seed: int, first_pair: int, count: int, L: int, threads: int = 1):
    lags = np.zeros(count, np.int64); coefs = np.zeros(count, np.float64)
    rets = np.zeros(count, np.int32); peaks = np.zeros(count, np.float64)
    seconds = np.zeros(count, np.float64)
    rc = lib().oracle_synth_batch(seed, first_pair, count, L, threads,
                                  lags.ctypes.data_as(C.POINTER(C.c_long)), _dptr(coefs),
                                  rets.ctypes.data_as(C.POINTER(C.c_int)), _dptr(peaks),
                                  _dptr(seconds))
    if rc != 0:
        raise MemoryError("oracle_synth_batch failed")
    return dict(lags=lags, coefs=coefs, rets=rets, peaks=peaks, seconds=seconds)
