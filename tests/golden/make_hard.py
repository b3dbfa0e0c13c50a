#!/usr/bin/env python3
"""Writes tests/golden/hard.json -- run in the authoring container only (needs oracle/_ref).

Low-margin and edge-case goldens at the interval-schedule lengths: for every
L of src/audiosync.c:50-57 three pairs whose two best lags differ by about 3e-4, 1e-3 and 1e-2 of
the peak (parameters tuned here by bisection on the NumPy oracle's margin), the extreme lags
0, +1, -1, L-1, the fold boundary idx == L and idx == L + 1, and the all-zero sample at
L = 144,000 and 1,440,000.  Recorded per case: what the compiled, unmodified reference
(oracle/_ref) returned -- (ret, lag, coef) -- plus raw index, peak, second peak, margin and
normalised correlation from the restatement.  Inputs are rebuilt from `params` by hard_cases.py.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import hard_cases as hc  # noqa: E402
from oracle import capi, xcorr_numpy as xn  # noqa: E402

TARGETS = (3e-4, 1e-3, 1e-2)


def margin_of(case):
    s, m = hc.build(case)
    o = xn.cross_correlation(s, m)
    return o["margin"], o


def tune(case, key, lo, hi, target, increasing):
    """Integer bisection of params[key] in [lo, hi] so that the oracle's margin lands next to target."""
    best = None
    while lo <= hi:
        mid = (lo + hi) // 2
        case["params"][key] = mid
        mg, o = margin_of(case)
        expect = case["params"]["d1"] % (2 * case["L"])
        ok = o["raw_index"] == expect
        if ok and (best is None or abs(np.log(mg / target)) < abs(np.log(best[1] / target))):
            best = (mid, mg)
        too_big = (mg > target) if ok else False          # peak lost: margin has gone below zero
        if too_big == increasing:
            hi = mid - 1
        else:
            lo = mid + 1
    assert best is not None, case
    case["params"][key] = best[0]
    return best[1]


def record(case):
    s, m = hc.build(case)
    L = case["L"]
    ret, lag, coef = capi.ref_cross_correlation(s, m)
    ex = capi.cross_correlation(s, m)
    assert (ex["ret"], ex["lag"]) == (ret, lag)
    assert ex["coef"] == coef or (coef != coef and ex["coef"] != ex["coef"])
    nn = lambda v: None if v != v else v
    case.update(ret=ret, lag=lag, coef=nn(coef), raw_index=ex["raw_index"], peak=ex["peak"], second=ex["second"],
                margin=nn(ex["margin"]), ncc=nn(ex["ncc"]), success=bool(ret == 0 and coef >= 0.95))
    print("%-12s L=%-8d %-40s idx=%-8d lag=%-8d ret=%2d coef=%s margin=%s" % (
        case["kind"], L, json.dumps(case["params"])[:40], ex["raw_index"], lag, ret,
        "nan" if coef != coef else "%.6f" % coef, "%.3e" % ex["margin"]), flush=True)
    return case


def main():
    capi.build()
    assert capi.ref_lib() is not None, "needs /root/reference to build oracle/_ref"
    cases = []
    for li, L in enumerate(xn.INTERV_SAMPLE):
        for ti, target in enumerate(TARGETS):
            seed = 0xA000 + 16 * li + ti
            neg = (li + ti) % 2 == 1
            if (li + ti) % 3 != 1:
                d1 = -(L // 5) - 7 if neg else L // 3 + 11
                # negative lags: the echo lies nearer to lag 0 (its overlap, hence its energy, is larger)
                d2 = d1 + L // 9 if neg else d1 + (L // 7 if (li % 2) else -(L // 9))
                c = dict(kind="echo", L=L, target=target, params=dict(seed=seed, d1=d1, d2=d2, k4096=4000))
                tune(c, "k4096", 2048, 4095, target, increasing=False)
            else:
                d1 = -(L // 4) + 3 if neg else L // 6 + 5
                c = dict(kind="periodic", L=L, target=target, params=dict(seed=seed, d1=d1, period=48 + 16 * ti, kn=20000))
                tune(c, "kn", 64, 1 << 19, target, increasing=True)
            cases.append(record(c))
    for L in (144000, 1440000):
        for lag in (0, 1, -1, L - 1):
            cases.append(record(dict(kind="lag", L=L, params=dict(seed=0xB000 + (lag % 97), lag=lag))))
        cases.append(record(dict(kind="impulse", L=L, note="idx == L: empty window, NaN, lag = -L", params=dict(i_src=L, i_smp=0))))
        cases.append(record(dict(kind="impulse", L=L, note="idx == L + 1: lag = -L + 1, one-frame window", params=dict(i_src=0, i_smp=L - 1))))
        cases.append(record(dict(kind="zero_sample", L=L, params=dict(seed=0xB100))))
    with open(os.path.join(HERE, "hard.json"), "w") as f:
        json.dump(dict(source="oracle/_ref (reference src/cross_correlation.c unmodified, backend %s); inputs: "
                              "tests/golden/hard_cases.py" % capi.backend(), cases=cases), f, indent=1)
    print("hard.json written:", len(cases), "cases")


if __name__ == "__main__":
    main()
