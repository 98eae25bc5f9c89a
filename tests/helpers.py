"""Shared helpers for the parity tests (golden loading, special-value parsing)."""
import json
import math
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fnum(x):
    if x is None:
        return None
    if isinstance(x, str):
        return {"nan": math.nan, "inf": math.inf, "-inf": -math.inf}[x]
    return float(x)


def farr(xs):
    return np.array([fnum(x) for x in xs], dtype=np.float64)


def load_unit():
    with open(os.path.join(GOLDEN_DIR, "reference_unit_vectors.json")) as f:
        return json.load(f)


def load_sqlness():
    with open(os.path.join(GOLDEN_DIR, "reference_sqlness_vectors.json")) as f:
        return json.load(f)


def udf_case_inputs(case, unit):
    if "fixture" in case:
        fx = unit["fixtures"][case["fixture"]]
        ts, val, ranges = fx["ts"], fx["val"], fx["ranges"]
    else:
        ts, val, ranges = case["ts"], case["val"], case["ranges"]
    return np.array(ts, np.int64), farr(val), np.array(ranges, np.uint32).reshape(-1, 2)


def check_expected(out, valid, expected, tol, what=""):
    assert len(out) == len(expected), what
    for i, e in enumerate(expected):
        if e is None:
            assert not valid[i], f"{what}: window {i} expected null, got {out[i]}"
            continue
        e = fnum(e)
        assert valid[i], f"{what}: window {i} expected {e}, got null"
        if math.isnan(e):
            assert math.isnan(out[i]), f"{what}: window {i}"
        elif tol == 0 or math.isinf(e):
            assert out[i] == e, f"{what}: window {i}: {out[i]!r} != {e!r}"
        else:
            assert abs(out[i] - e) < tol, f"{what}: window {i}: {out[i]!r} vs {e!r}"


def pack_series(series_dict):
    """{name: {ts, val}} -> (names, ts, val, sid, offsets) in dict order (series sorted rows)."""
    names, ts, val, sid, offsets = [], [], [], [], [0]
    for i, (name, d) in enumerate(series_dict.items()):
        names.append(name)
        ts.extend(d["ts"])
        val.extend(fnum(v) for v in d["val"])
        sid.extend([i] * len(d["ts"]))
        offsets.append(len(ts))
    return (names, np.array(ts, np.int64), np.array(val, np.float64), np.array(sid, np.uint32),
            np.array(offsets, np.uint64))


def promql_series(spec):
    """Prometheus test-series notation 'a+bxN a-bxN ...' (double_exponential_smoothing.rs:466-494)."""
    out = []
    for part in spec.split(" "):
        head, n = part.split("x")
        n = int(n)
        if "+" in head:
            a, b = head.split("+")
            a, b = int(a), int(b)
            out.extend(float(x) for x in range(a, b * n + a + 1, b))
        else:
            a, b = head.split("-")
            a, b = int(a), int(b)
            lo = -b * n + a
            seq = list(range(lo, a + 1))[::-1][::b]
            out.extend(float(x) for x in seq)
    return np.array(out, np.float64)
