#!/usr/bin/env python3
"""Regenerates tests/golden/*.json -- run in the authoring container only.

Sources of truth, in order:
  * kat.json       -- the 12 known-answer checks transcribed from the
                      reference's tests/test_cross_correlation.c:13-116 and
                      tests/test_pearson_coefficient.c:13-61 (inputs and the
                      asserted outcomes), plus the values the reference's own
                      compiled code (oracle/_ref, FFT shim underneath) returned
                      for them when this script ran.
  * synth.json     -- (ret, lag, coef) returned by oracle/_ref's compiled,
                      unmodified reference cross_correlation() on seeded
                      synthetic pairs (generator: SURVEY.md 8d /
                      oracle/xcorr_oracle.c), plus peak / second-peak from the
                      restatement so tests can apply the "unique peak" rule.
/root/reference is needed to (re)build oracle/_ref; the JSON files are what
travels to the GPU box.
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import capi, xcorr_numpy as xn  # noqa: E402

SEED = 0x5EED


def kat_cases():
    sin7_src = [math.sin(i) for i in range(2000)]
    sin7_smp = [math.sin(i) for i in range(1000)]
    sin8_src = [math.sin(i + 180) for i in range(1000)] + [0.0] * 1000
    sin8_smp = [math.sin(i) for i in range(1000)]
    xc = [
        # name, source, sample, expectations as asserted by the reference test
        ("T1", [1.1, 2.2, 3.3, 4.4, 5.5, 0, 0, 0, 0, 0], [1.1, 2.2, 3.3, 4.4, 5.5],
         dict(ret=0, lag=0, coef_eq=1.0)),
        ("T2", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14], [0, 0, 0, 0, 0, 0, 0],
         dict(ret=-1)),
        ("T3", [0, 0, 0, 1, 2, 3, 4, 5, 6, 0, 0, 0], [1, 2, 3, 4, 5, 6],
         dict(ret=0, lag=3, coef_gt=0.95)),
        ("T4", [1, 2, 3, 0.4, 1.1, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 2, 3],
         dict(ret=0, lag=-3, coef_gt=0.95)),
        ("T5", [1, 2, 3, 4, -1.0, 0, 0, 4, 3, 2, 1, 0, 0, 0], [0, 0, 0, 1, 2, 3, 4],
         dict(ret=0, lag=-3, coef_gt=0.95)),
        ("T6", [0, 0, 0, 0, 0, 1, 2, 3, 4, -1, -3, -5, 0, 0], [1, 2, 3, 4, -1, -3, -5],
         dict(ret=0, lag=5, coef_gt=0.95)),
        ("T7", "sin(i), i<2000", "sin(i), i<1000", dict(ret=0, lag=0, coef_gt=0.95)),
        ("T8", "sin(i+180) i<1000 then 1000 zeros", "sin(i), i<1000",
         dict(ret=0, lag=-1, coef_lt=-0.95)),
    ]
    out = []
    for name, src, smp, exp in xc:
        if name == "T7":
            s, p = sin7_src, sin7_smp
        elif name == "T8":
            s, p = sin8_src, sin8_smp
        else:
            s, p = [float(v) for v in src], [float(v) for v in smp]
        ret, lag, coef = capi.ref_cross_correlation(np.array(s), np.array(p))
        rec = dict(name=name, expect=exp, ref=dict(ret=ret, lag=lag,
                                                   coef=None if coef != coef else coef))
        if name in ("T7", "T8"):
            rec["generator"] = dict(source=src, sample=smp)
        else:
            rec["source"], rec["sample"] = s, p
        out.append(rec)
    # pearson: (x window, y window, expectation); windows already resolved
    s1 = [1.0, 2.1, 3.2, 4.3, 5.4, 6.5, 7.6, 8.7, 9.8, 10.9]
    p1 = [0, 0, 1.0, 2.1, 3.2, 4.3, 5.4, 6.5, 7.6, 8.7]
    s2 = [0, 0, 0, 0, 100, 200, 300, 400, 500, 600, 700]
    p2 = [100, 200, 300, 400, 500]
    pe = [
        ("P1", s1[0:8], p1[2:10], dict(eq=1.0)),       # lag = -2, len 10
        ("P2", s2[4:9], p2[0:5], dict(eq=1.0)),        # lag = +4, len 5
        ("P3", [1, 2, 3, 4], [4, 3, 2, 1], dict(eq=-1.0)),
        ("P4", [1, 2, 3, 4], [0, 0, 0, 0], dict(nan=True)),
    ]
    pout = []
    for name, x, y, exp in pe:
        v = capi.ref_pearson(np.array(x, float), np.array(y, float))
        pout.append(dict(name=name, x=[float(a) for a in x], y=[float(a) for a in y],
                         expect=exp, ref=None if v != v else v))
    return dict(cross_correlation=out, pearson=pout)


def synth_cases():
    cases = []

    def add(L, pid, source=None, sample=None, tag="pair", seed=SEED):
        if source is None:
            source, sample = capi.synth_pair(SEED, pid, L)
        ret, lag, coef = capi.ref_cross_correlation(source[:2 * L], sample[:L])
        ex = capi.cross_correlation(source[:2 * L], sample[:L])
        assert (ex["ret"], ex["lag"]) == (ret, lag)
        assert (ex["coef"] == coef) or (coef != coef and ex["coef"] != ex["coef"])
        margin = (abs(ex["peak"]) - ex["second"]) / abs(ex["peak"]) if ex["peak"] != 0 else 0.0
        cases.append(dict(tag=tag, seed=seed, pair_id=pid, L=L, ret=ret, lag=lag,
                          coef=None if coef != coef else coef, peak=ex["peak"],
                          second=ex["second"], margin=margin,
                          raw_index=ex["raw_index"],
                          true_lag=capi.synth_true_lag(SEED, pid, L) if tag == "pair" else None,
                          success=bool(ret == 0 and coef >= 0.95)))

    for L in (6, 50, 64, 250, 1000, 1024, 3600, 6000, 24000):
        for pid in (0, 1, 2, 3):
            add(L, pid)
    for L in xn.INTERV_SAMPLE:
        for pid in (0, 3, 5):
            add(L, pid)
    for pid in (1, 2, 7, 11):
        add(1440000, pid)
    # config 2: one full-length pair evaluated on the interval prefixes
    for pid in (0, 6):
        src, smp = capi.synth_pair(SEED + 2, pid, 1440000)
        for L in xn.INTERV_SAMPLE:
            add(L, pid, src, smp, tag="interval-prefix", seed=SEED + 2)
    return cases


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    capi.build()
    assert capi.ref_lib() is not None, "needs /root/reference to build oracle/_ref"
    with open(os.path.join(here, "kat.json"), "w") as f:
        json.dump(dict(source="reference tests/test_cross_correlation.c + "
                              "tests/test_pearson_coefficient.c; ref values from oracle/_ref "
                              "(backend %s)" % capi.backend(), **kat_cases()), f, indent=1)
    with open(os.path.join(here, "synth.json"), "w") as f:
        json.dump(dict(source="oracle/_ref (reference src/cross_correlation.c unmodified, "
                              "backend %s)" % capi.backend(), cases=synth_cases()), f, indent=1)
    print("golden written")


if __name__ == "__main__":
    main()
