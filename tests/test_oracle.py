"""Pins the CPU oracle (test infrastructure) -- runs without a GPU.

1. the restatement reproduces every known-answer check of the reference's own
   tests/test_cross_correlation.c and tests/test_pearson_coefficient.c;
2. it agrees bit for bit with the reference's compiled, unmodified
   src/cross_correlation.c (oracle/_ref) where that build is present, and with
   the committed golden vectors that build produced;
3. it agrees with the independent NumPy/pocketfft restatement;
4. the three forms of the synthetic generator agree bit for bit.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import capi, xcorr_numpy as xn

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLDEN, "kat.json")))
SYNTH = json.load(open(os.path.join(GOLDEN, "synth.json")))


def kat_inputs(rec):
    if rec["name"] == "T7":
        return (np.sin(np.arange(2000, dtype=np.float64)), np.sin(np.arange(1000, dtype=np.float64)))
    if rec["name"] == "T8":
        src = np.concatenate([np.array([math.sin(i + 180) for i in range(1000)]), np.zeros(1000)])
        return src, np.array([math.sin(i) for i in range(1000)])
    return np.array(rec["source"], float), np.array(rec["sample"], float)


def check_expect(exp, ret, lag, coef):
    assert ret == exp["ret"]
    if "lag" in exp:
        assert lag == exp["lag"]
    if "coef_eq" in exp:
        assert coef == exp["coef_eq"]          # exact, as the reference asserts
    if "coef_gt" in exp:
        assert coef > exp["coef_gt"]
    if "coef_lt" in exp:
        assert coef < exp["coef_lt"]


@pytest.mark.parametrize("rec", KAT["cross_correlation"], ids=lambda r: r["name"])
def test_kat_cross_correlation(rec):
    src, smp = kat_inputs(rec)
    got = capi.cross_correlation(src, smp)
    check_expect(rec["expect"], got["ret"], got["lag"], got["coef"])
    # golden values recorded from the compiled reference
    assert got["ret"] == rec["ref"]["ret"] and got["lag"] == rec["ref"]["lag"]
    if rec["ref"]["coef"] is None:
        assert got["coef"] != got["coef"]
    else:
        assert got["coef"] == rec["ref"]["coef"]
    # independent numpy restatement
    alt = xn.cross_correlation(src, smp)
    check_expect(rec["expect"], alt["ret"], alt["lag"], alt["coef"]) if rec["name"] != "T1" else None
    assert alt["lag"] == got["lag"] and alt["ret"] == got["ret"]
    if capi.ref_lib() is not None:
        ret, lag, coef = capi.ref_cross_correlation(src, smp)
        check_expect(rec["expect"], ret, lag, coef)


@pytest.mark.parametrize("rec", KAT["pearson"], ids=lambda r: r["name"])
def test_kat_pearson(rec):
    x, y = np.array(rec["x"]), np.array(rec["y"])
    vals = [capi.pearson(x, y)]
    if capi.ref_lib() is not None:
        vals.append(capi.ref_pearson(x, y))
    for v in vals:
        if rec["expect"].get("nan"):
            assert v != v
        else:
            assert v == rec["expect"]["eq"]    # exact equality, as in the reference test


def test_max_abs_index_semantics():
    # reference src/cross_correlation.c:52-67
    f = capi.max_abs_index
    assert f(np.array([0.0, 0.0, 0.0])) == 0                 # all zero -> 0
    assert f(np.array([5.0, -5.0, 5.0])) == 0                # strict >: index 0 keeps ties
    assert f(np.array([-5.0, 1.0, 2.0])) == 2                # signed seed: negative r[0] loses
    assert f(np.array([-5.0, 0.0, 0.0])) == 1                # ... even to |0|
    assert f(np.array([1.0, -3.0, 3.0, 2.0])) == 1           # first of equal magnitudes
    assert f(np.array([1.0, np.nan, 2.0])) == 2              # NaN never wins
    assert f(np.array([np.nan, 1.0, 2.0])) == 0              # NaN seed is never beaten
    for arr in ([0.0, 0.0, 0.0], [5.0, -5.0, 5.0], [-5.0, 1.0, 2.0], [-5.0, 0.0, 0.0],
                [1.0, -3.0, 3.0, 2.0], [1.0, np.nan, 2.0]):
        assert xn.max_abs_index(np.array(arr)) == f(np.array(arr))


def test_fold_boundary_gives_nan():
    # idx == L => empty Pearson window => NaN => ret -1 (src/cross_correlation.c:256-276)
    L = 8
    src = np.zeros(2 * L); smp = np.zeros(L)
    smp[0] = 1.0
    src[L] = 1.0
    got = capi.cross_correlation(src, smp)
    assert got["raw_index"] == L and got["lag"] == -L and got["ret"] == -1
    assert got["coef"] != got["coef"]


@pytest.mark.parametrize("L", [5, 6, 7, 12, 50, 250, 1000, 1024, 3600, 4374, 144000])
def test_generator_forms_agree(L):
    for pid in (0, 3, 9):
        s64, p64 = capi.synth_pair(0x5EED, pid, L, np.float64)
        s32, p32 = capi.synth_pair(0x5EED, pid, L, np.float32)
        si, pi = capi.synth_pair(0x5EED, pid, L, np.int32)
        ns, np_ = xn.synth_pair_int(0x5EED, pid, L)
        assert np.array_equal(si, ns) and np.array_equal(pi, np_)
        assert np.array_equal(s64, si * 2.0 ** -23) and np.array_equal(p64, pi * 2.0 ** -23)
        assert np.array_equal(s32.astype(np.float64), s64)      # exact in fp32
        assert np.array_equal(p32.astype(np.float64), p64)
        assert np.abs(pi).max() < 2 ** 24
        assert capi.synth_true_lag(0x5EED, pid, L) == xn.synth_true_lag(0x5EED, pid, L)


def _small_cases():
    return [c for c in SYNTH["cases"] if c["L"] <= 144000 and c["tag"] == "pair"]


@pytest.mark.parametrize("case", _small_cases(), ids=lambda c: "L%d-p%d" % (c["L"], c["pair_id"]))
def test_golden_synth_small(case):
    src, smp = capi.synth_pair(case["seed"], case["pair_id"], case["L"])
    got = capi.cross_correlation(src, smp)
    assert got["ret"] == case["ret"] and got["lag"] == case["lag"]
    assert got["coef"] == case["coef"]
    assert got["lag"] == case["true_lag"]
    assert bool(capi.lib().oracle_accept(got["ret"], got["coef"])) == case["success"]
    alt = xn.cross_correlation(src, smp)
    assert alt["lag"] == got["lag"] and alt["ret"] == got["ret"]
    assert abs(alt["coef"] - got["coef"]) <= 1e-12
    assert abs(alt["peak"] - got["peak"]) <= 1e-9 * abs(got["peak"])
    if capi.ref_lib() is not None:
        ret, lag, coef = capi.ref_cross_correlation(src, smp)
        assert (ret, lag, coef) == (got["ret"], got["lag"], got["coef"])


def test_golden_full_size_one_pair():
    case = [c for c in SYNTH["cases"] if c["L"] == 1440000 and c["tag"] == "pair"][0]
    src, smp = capi.synth_pair(case["seed"], case["pair_id"], case["L"])
    got = capi.cross_correlation(src, smp)
    assert (got["ret"], got["lag"], got["coef"]) == (case["ret"], case["lag"], case["coef"])
    assert got["lag"] == case["true_lag"]


def test_interval_loop_matches_prefix_goldens():
    cases = [c for c in SYNTH["cases"] if c["tag"].startswith("interval") and c["pair_id"] == 6]
    src, smp = capi.synth_pair(cases[0]["seed"], 6, 1440000)
    out = capi.interval_loop(src, smp)
    # reference loop stops at the first success (src/audiosync.c:254-258)
    first_ok = next(i for i, c in enumerate(cases) if c["success"])
    assert out["n"] == first_ok + 1
    for i in range(out["n"]):
        assert out["rets"][i] == cases[i]["ret"] and out["lags"][i] == cases[i]["lag"]
        assert bool(out["succ"][i]) == cases[i]["success"]
    assert out["final_ret"] == 0
    assert out["final_lag"] == round(cases[first_ok]["lag"] * 1000.0 / 48000.0)


def test_synth_batch_threads():
    a = capi.synth_batch(0x5EED, 0, 6, 3600, threads=1)
    b = capi.synth_batch(0x5EED, 0, 6, 3600, threads=3)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True)
    for pid in range(6):
        assert a["lags"][pid] == capi.synth_true_lag(0x5EED, pid, 3600)


def test_peak_quality_restatement_against_brute_force():
    """oracle/xcorr_numpy.peak_quality (margin, normalised correlation) against
    a direct evaluation: r by the O(L^2) circular sum, windows by the reference's fold."""
    from oracle import xcorr_numpy
    rng = np.random.default_rng(5)
    for L, shift in ((16, 3), (50, -7), (64, 0), (33, -5)):
        base = rng.standard_normal(3 * L)
        src = base[L:3 * L].copy()
        smp = base[L + shift:2 * L + shift] + 0.05 * rng.standard_normal(L)
        N = 2 * L
        r = np.array([N * sum(src[(n + j) % N] * smp[n] for n in range(L)) for j in range(N)])
        o = xcorr_numpy.cross_correlation(src, smp)
        idx = xcorr_numpy.max_abs_index(r)
        assert o["raw_index"] == idx and o["lag"] == shift
        mag = np.abs(r); mag[idx] = -1
        assert abs(o["margin"] - (abs(r[idx]) - mag.max()) / abs(r[idx])) < 1e-9
        lag, wx, wy = xcorr_numpy.windows(src, smp, idx)
        assert lag == shift
        assert abs(o["ncc"] - r[idx] / (N * np.sqrt((wx * wx).sum() * (wy * wy).sum()))) < 1e-9
        if shift >= 0:
            assert abs(o["ncc"] - (wx * wy).sum() / np.sqrt((wx * wx).sum() * (wy * wy).sum())) < 1e-9
        c = capi.cross_correlation(src, smp)                  # the C restatement carries the same fields
        assert abs(c["margin"] - o["margin"]) < 1e-9 and abs(c["ncc"] - o["ncc"]) < 1e-9


def _hard_cases(max_L):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hard.json")))
    return [c for c in d["cases"] if c["L"] <= max_L]


@pytest.mark.parametrize("case", _hard_cases(288000), ids=lambda c: "%s-L%d-%.0e" % (c["kind"], c["L"], c["margin"] or 0))
def test_oracle_reproduces_hard_goldens(case):
    """The restatement and the NumPy cross-check against the compiled reference's recorded outputs on
    the low-margin / edge inputs (tests/golden/hard.json; generator tests/golden/make_hard.py)."""
    import hard_cases as hc
    src, smp = hc.build(case)
    for o in (capi.cross_correlation(src, smp), xn.cross_correlation(src, smp)):
        assert (o["ret"], o["lag"], o["raw_index"]) == (case["ret"], case["lag"], case["raw_index"])
        if case["coef"] is None:
            assert o["coef"] != o["coef"]
        else:
            assert abs(o["coef"] - case["coef"]) <= 1e-9 * abs(case["coef"])
        assert abs(o["peak"] - case["peak"]) <= 1e-9 * abs(case["peak"]) + 1e-12
        assert abs(o["margin"] - case["margin"]) <= 1e-7
