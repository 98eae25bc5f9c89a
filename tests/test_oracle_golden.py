"""Pins the CPU oracle against the reference's own known-answer vectors (CPU only).

Every case in tests/golden/*.json cites the reference #[test] / sqlness file it was ported from.
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import (check_expected, farr, fnum, load_sqlness, load_unit, pack_series, promql_series,
                           udf_case_inputs)

UNIT = load_unit()
SQL = load_sqlness()


@pytest.mark.parametrize("case", UNIT["range_udf"], ids=lambda c: c["name"])
def test_range_udf_golden(case):
    ts, val, ranges = udf_case_inputs(case, UNIT)
    out, valid = orc.range_udf(case["fn"], ts, val, ranges, eval_ts=case.get("eval_ts"),
                               range_length=case.get("range_length", 0), param0=case.get("param0", 0.0),
                               param1=case.get("param1", 0.0))
    check_expected(out, valid, case["expected"], case["tol"], case["name"])
    # nulls leave 0.0 in the raw buffer like an Arrow builder (extrapolate_rate.rs:497-523 asserts it)
    assert all(out[i] == 0.0 for i, e in enumerate(case["expected"]) if e is None)


@pytest.mark.parametrize("case", [c for c in UNIT["range_udf"] if c["fn"] in ("rate", "increase")],
                         ids=lambda c: c["name"])
def test_rate_rescan_equals_sliding_on_goldens(case):
    ts, val, ranges = udf_case_inputs(case, UNIT)
    a, va = orc.range_udf(case["fn"], ts, val, ranges, case["eval_ts"], case["range_length"])
    b, vb = orc.range_udf(case["fn"], ts, val, ranges, case["eval_ts"], case["range_length"], rescan=True)
    assert (va == vb).all() and (a == b).all()


def test_holt_winters_trends():
    g = UNIT["holt_winters_trends"]
    for spec, expected in g["cases"]:
        v = promql_series(spec)
        ts = np.arange(v.size, dtype=np.int64)
        out, valid = orc.range_udf("holt_winters", ts, v, [[0, 801]], param0=0.01, param1=0.1)
        assert valid[0] and abs(out[0] - expected) < 1e-4, (spec, out[0])


def test_holt_winters_impl():
    for c in UNIT["holt_winters_impl"]["cases"]:
        r = orc.holt_winters(farr(c["values"]), fnum(c["sf"]), fnum(c["tf"]))
        e = fnum(c["expected"])
        assert (math.isnan(r) and math.isnan(e)) or r == e, c


def test_quantile_impl():
    for c in UNIT["quantile_impl"]["cases"]:
        r = orc.quantile(farr(c["values"]), fnum(c["q"]))
        e = fnum(c["expected"])
        assert (math.isnan(r) and math.isnan(e)) or r == e, c


def test_linear_regression():
    for c in UNIT["linear_regression"]["cases"]:
        n = len(c["val"])
        s, i = orc.linear_regression(c["ts"][:n], farr(c["val"]), c["intercept_time"])
        assert s == c["slope"] and i == c["intercept"], c["name"]
    k = UNIT["linear_regression"]["kahan"]
    s, c = orc.compensated_sum(k["inputs"])
    assert s + c == k["expected_sum_plus_c"]


def test_histogram_evaluate_row():
    for c in UNIT["histogram_evaluate_row"]["cases"]:
        v, err = orc.histogram_evaluate_row(c["q"], farr(c["bucket"]), farr(c["counters"]))
        if c["expected"] == "err":
            assert err
            continue
        assert not err
        e = fnum(c["expected"])
        if math.isnan(e):
            assert math.isnan(v), c
        elif "tol" in c:
            assert abs(v - e) < c["tol"], c
        else:
            # the reference compares format!("{actual}") == format!("{expected}")
            assert repr(v) == repr(e), (c, v)


def test_range_manipulate():
    g = UNIT["range_manipulate"]
    for c in g["cases"]:
        for definitional in (False, True):
            off, ln, s2, e2 = orc.calculate_range(g["ts"], c["start"], c["end"], c["interval"], c["range"], definitional)
            assert list(zip(off.tolist(), ln.tolist())) == [tuple(r) for r in c["ranges"]], c["name"]
            assert list(range(s2, e2 + 1, c["interval"])) == c["eval_ts"], c["name"]
    a = g["alignment"]
    off, ln, s2, e2 = orc.calculate_range(a["ts"], a["start"], a["end"], a["interval"], a["range"])
    assert s2 % a["interval"] == a["start"] % a["interval"]
    assert len(off) == len(range(s2, e2 + 1, a["interval"]))


def test_range_manipulate_cursor_overshoot_quirk():
    """the literal cursor walk reports an empty window when its cursor overshoots
    len, although samples lie inside (t-range, t]; the definitional variant finds them."""
    ts = [0, 1, 2, 3, 4, 100, 149, 150]
    off, ln, _, _ = orc.calculate_range(ts, 0, 150, 50, 10)
    offd, lnd, _, _ = orc.calculate_range(ts, 0, 150, 50, 10, definitional=True)
    assert list(zip(offd.tolist(), lnd.tolist()))[-1] == (6, 2)
    assert list(zip(off.tolist(), ln.tolist()))[-1] == (0, 0)


def test_instant_manipulate():
    g = UNIT["instant_manipulate"]
    for c in g["cases"]:
        d = g["data_nan"] if c["nan"] else g["data"]
        take, ots = orc.instant_manipulate(d["ts"], farr(d["val"]), c["start"], c["end"], c["interval"], c["lookback"])
        assert ots.tolist() == c["out_ts"], c["name"]
        if "out_val" in c:
            assert farr(d["val"])[take.astype(np.int64)].tolist() == c["out_val"], c["name"]


def test_normalize():
    g = UNIT["normalize"]
    for c in g["cases"]:
        ts, val = orc.normalize(g["ts"], farr(g["val"]), c["offset"], c["filter_nan"])
        assert ts.tolist() == c["out_ts"] and val.tolist() == c["out_val"], c["name"]
    ts, val = orc.normalize([1, 2, 3], farr([1.0, "nan", 3.0]), 10, True)
    assert ts.tolist() == [11, 13] and val.tolist() == [1.0, 3.0]
    ts, val = orc.normalize([1, 2, 3], farr([1.0, "nan", 3.0]), 10, False)
    assert ts.tolist() == [11, 12, 13]


def test_series_divide():
    offs = orc.series_divide([5, 5, 5, 7, 7, 9])
    assert offs.tolist() == [0, 3, 5, 6]
    assert orc.series_divide([]).tolist() == [0]
    assert orc.series_divide([1]).tolist() == [0, 1]


def _series_divide_ids(g):
    """Dense ids by key change over the concatenated batches (what the plan layer hands K0)."""
    keys = [tuple(b[t][i] for t in g["tag_columns"]) for b in g["batches"] for i in range(len(b["ts"]))]
    ids, cur = [], -1
    for i, k in enumerate(keys):
        if i == 0 or k != keys[i - 1]:
            cur += 1
        ids.append(cur)
    return keys, np.array(ids, np.uint32)


def test_series_divide_reference_table():
    """The reference's own SeriesDivide fixture (series_divide.rs:668-905: three input batches, a series that spans all
    of them, multi-byte tag values): 7 series with the row counts of `per_batch_data`."""
    g = UNIT["series_divide"]
    keys, ids = _series_divide_ids(g)
    assert orc.series_divide(ids).tolist() == g["expected_offsets"]
    firsts = [keys[o] for o in g["expected_offsets"][:-1]]
    assert firsts == [(e["host"], e["path"]) for e in g["expected_series"]]
    assert np.diff(g["expected_offsets"]).tolist() == [e["rows"] for e in g["expected_series"]]


def _resolve_series(case):
    return case["series"] if "series" in case else SQL["series_sets"][case["series_ref"]]


@pytest.mark.parametrize("mode", ["flat", "faithful"])
@pytest.mark.parametrize("case", SQL["range_cases"], ids=lambda c: c["name"])
def test_sqlness_range_cases(case, mode):
    names, ts, val, sid, offsets = pack_series(_resolve_series(case))
    p = orc.make_params(case["fn"], case["start"], case["end"], case["interval"], case["range"],
                        offset=case["offset"], filter_nan=True, param0=case.get("param0", 0.0))
    out, valid = orc.range_query(p, ts, val, sid, offsets, mode=mode, threads=2)
    T = orc.num_steps(case["start"], case["end"], case["interval"])
    vb = orc.valid_to_bool(valid, T)
    got = {}
    for s, name in enumerate(names):
        for k in range(T):
            if vb[s, k]:
                got[(name, case["start"] + k * case["interval"])] = out[s, k]
    exp = {(n, t): fnum(v) for n, t, v in case["expected"]}
    assert set(got) == set(exp), (case["name"], got)
    for key, e in exp.items():
        assert got[key] == e or (math.isnan(e) and math.isnan(got[key])), (case["name"], key, got[key], e)


@pytest.mark.parametrize("case", SQL["instant_cases"], ids=lambda c: c["name"])
def test_sqlness_instant_cases(case):
    names, ts, val, sid, offsets = pack_series(case["series"])
    out, valid = orc.instant_query(ts, val, offsets, case["start"], case["end"], case["interval"], case["lookback"],
                                   case["offset"])
    T = orc.num_steps(case["start"], case["end"], case["interval"])
    vb = orc.valid_to_bool(valid, T)
    got = {(names[s], case["start"] + k * case["interval"]): out[s, k]
           for s in range(len(names)) for k in range(T) if vb[s, k]}
    exp = {(n, t): fnum(v) for n, t, v in case["expected"]}
    assert got == exp


@pytest.mark.parametrize("case", SQL.get("instant_offset_direction_cases", []), ids=lambda c: c["name"])
def test_sqlness_instant_offset_direction(case):
    """promql/offset_direction.result: a positive offset reads the past, a negative one the future."""
    test_sqlness_instant_cases(case)


@pytest.mark.parametrize("case", SQL["histogram_cases"], ids=lambda c: c["name"])
def test_sqlness_histogram_cases(case):
    B = len(case["le"])
    series = {f"b{b}": {"ts": case["bucket_ts"], "val": case["bucket_val"][b]} for b in range(B)}
    names, ts, val, sid, offsets = pack_series(series)
    p = orc.make_params(case["fn"], case["start"], case["end"], case["interval"], case["range"], offset=case["offset"])
    rates, valid = orc.range_query(p, ts, val, sid, offsets)
    T = orc.num_steps(case["start"], case["end"], case["interval"])
    # `sum by (le, s)` with one series per group is the identity on values (planner.rs:334-452)
    for q, expected in case["quantiles"]:
        out, ov = orc.histogram_quantile(q, farr(case["le"]), rates, valid)
        vb = orc.valid_to_bool(ov, T)
        exp = expected if isinstance(expected, list) else [expected]
        got = [out[0, k] for k in range(T) if vb[0, k]]
        assert got == exp, (case["name"], q, got)


def test_group_aggregate_matches_numpy():
    rng = np.random.default_rng(7)
    S, T, G = 64, 40, 5
    vals = rng.normal(size=(S, T))
    vb = rng.random((S, T)) > 0.3
    valid = np.packbits(np.pad(vb, ((0, 0), (0, (-T) % 32))), axis=1, bitorder="little").view(np.uint32)
    gid = rng.integers(0, G, S).astype(np.uint32)
    s, c = orc.group_aggregate("sum", vals, valid, gid, G)
    for g in range(G):
        m = (gid == g)[:, None] & vb
        assert (c[g] == m.sum(0)).all()
        ref = np.zeros(T)
        for si in range(S):  # sequential series order, like one DataFusion partition
            if gid[si] == g:
                ref += np.where(vb[si], vals[si], 0.0)
        assert np.allclose(s[g], ref, rtol=0, atol=0) or (s[g] == ref).all()
    a, _ = orc.group_aggregate("avg", vals, valid, gid, G)
    assert np.allclose(a[c > 0], (s / np.maximum(c, 1))[c > 0])


def test_synth_generator_matches_reference_bench_recurrence():
    """closed-form generator == sequential recurrences of benches/bench_range_fn.rs:60-82 (scaled)."""
    n = 300
    for s in (0, 1, 5, 36, 37, 1000):
        ts, val, sid = orc.synth_fill(s, 1, n, 1_700_000_000_000, 15000, 1000, 0, 0x5EED)
        cur, ref = 0.0, []
        for i in range(n):
            cur += 1.0 + (i % 7) * 0.25
            ref.append(cur * (1 + s % 13))
        assert val.tolist() == ref
        assert (np.diff(ts) > 0).all() and ((ts - 1_700_000_000_000 - np.arange(n) * 15000) < 1000).all()
        ts, val, sid = orc.synth_fill(s, 1, n, 0, 15000, 0, 1, 0x5EED)
        cur, ref = 0.0, []
        for i in range(n):
            if i > 0 and (i + s) % 37 == 0:
                cur = 1.0
            else:
                cur += 1.0 + (i % 5) * 0.5
            ref.append(cur * (1 + s % 13))
        assert val.tolist() == ref
        assert (ts == np.arange(n) * 15000).all()


# ---------------------------------------------------------------------------------------------------
# by-label aggregators: the reference's own result tables (tests-integration/src/tests/promql_test.rs:343-665)
# ---------------------------------------------------------------------------------------------------
def _load_aggr():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_aggregator_vectors.json")) as f:
        return json.load(f)


AGGR = _load_aggr()


def aggregator_case_inputs(case):
    """series passing the label matcher, sorted by tag tuple like the scan; dense group ids of the by-labels."""
    sel = [s for s in AGGR["series"] if all(s[k] == v for k, v in case["filter"].items())]
    sel.sort(key=lambda s: tuple(s[t] for t in AGGR["tags"]))
    keys = sorted({tuple(s[b] for b in case["by"]) for s in sel})
    gid = np.array([keys.index(tuple(s[b] for b in case["by"])) for s in sel], np.uint32)
    ts = np.concatenate([np.array(s["ts"], np.int64) for s in sel])
    val = np.concatenate([np.array(s["val"], np.float64) for s in sel])
    offsets = np.concatenate([[0], np.cumsum([len(s["ts"]) for s in sel])]).astype(np.uint64)
    return sel, keys, gid, ts, val, offsets


def check_aggregator_rows(case, keys, out, cnt, ets):
    got = {}
    for g, key in enumerate(keys):
        for k in range(out.shape[1]):
            if cnt[g, k] > 0:
                got[(key, int(ets[k]))] = float(out[g, k])
    exp = {(tuple(lbl[b] for b in case["by"]), int(t)): float(v) for lbl, t, v in case["expected"]}
    assert set(got) == set(exp), (case["name"], sorted(got), sorted(exp))
    tol = case.get("rel_tol", 0.0)
    for k in exp:
        assert abs(got[k] - exp[k]) <= tol * abs(exp[k]), (case["name"], k, got[k], exp[k])


@pytest.mark.parametrize("case", AGGR["cases"], ids=lambda c: c["name"])
def test_by_label_aggregators_reference_tables(case):
    """InstantManipulate (lookback 0) + the by-label aggregate of the oracle == the reference's result tables: pins
    orc_group_aggregate (the restatement of DataFusion's sum / avg / count / min / max / stddev_pop / var_pop
    accumulators) on groups with several members."""
    sel, keys, gid, ts, val, offsets = aggregator_case_inputs(case)
    out, valid = orc.instant_query(ts, val, offsets, AGGR["start"], AGGR["end"], AGGR["interval"], AGGR["lookback"])
    ets = AGGR["start"] + AGGR["interval"] * np.arange(out.shape[1])
    agg, cnt = orc.group_aggregate(case["agg"], out, valid, gid, len(keys))
    check_aggregator_rows(case, keys, agg, cnt, ets)


def test_by_label_aggregators_combined_expressions():
    by_name = {c["name"]: c for c in AGGR["cases"]}
    vals = {}
    for name in ("combined_sum_by_job", "combined_min_by_job", "combined_max_by_job", "combined_avg_by_job"):
        case = by_name[name]
        sel, keys, gid, ts, val, offsets = aggregator_case_inputs(case)
        out, valid = orc.instant_query(ts, val, offsets, AGGR["start"], AGGR["end"], AGGR["interval"], AGGR["lookback"])
        agg, cnt = orc.group_aggregate(case["agg"], out, valid, gid, len(keys))
        vals[case["agg"]] = {k[0]: float(agg[g, 0]) for g, k in enumerate(keys)}
    four, two = AGGR["combined_checks"]
    for job, exp in four["expected"].items():
        assert vals["sum"][job] + vals["min"][job] + vals["max"][job] + vals["avg"][job] == exp
    for job, exp in two["expected"].items():
        assert vals["sum"][job] + vals["min"][job] == exp


def test_min_max_by_label_follow_the_total_order_on_nan_members():
    """arrow-rs aggregate min / max and DataFusion's MinMax accumulators compare floats with f64::total_cmp: a (positive)
    NaN is the greatest value.  max by (..) of a group with a NaN member is NaN, min by (..) ignores it."""
    vals = np.array([[1.0], [np.nan], [3.0], [2.0]])
    valid = np.ones((4, 1), np.uint32)
    gid = np.zeros(4, np.uint32)
    mx, c = orc.group_aggregate("max", vals, valid, gid, 1)
    mn, _ = orc.group_aggregate("min", vals, valid, gid, 1)
    assert c[0, 0] == 4 and np.isnan(mx[0, 0]) and mn[0, 0] == 1.0
    neg_nan = np.frombuffer(np.uint64(0xfff8000000000000).tobytes(), np.float64)[0]
    vals[1, 0] = neg_nan                       # -NaN sorts below everything
    mx, _ = orc.group_aggregate("max", vals, valid, gid, 1)
    mn, _ = orc.group_aggregate("min", vals, valid, gid, 1)
    assert mx[0, 0] == 3.0 and np.isnan(mn[0, 0])


# ---------------------------------------------------------------------------------------------------
# HistogramFold operator tests (histogram_fold.rs:1452-1630) and the le-label parse
# ---------------------------------------------------------------------------------------------------
def _load_fold():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_histogram_fold_vectors.json")) as f:
        return json.load(f)


FOLD = _load_fold()


def _fnum(x):
    return float("nan") if x == "NaN" else float(x)


@pytest.mark.parametrize("case", FOLD["cases"], ids=lambda c: c["name"])
def test_histogram_fold_operator_goldens(case):
    rows = [((r[0],), r[1], r[2], r[3]) for r in case["rows"]]
    got = orc.histogram_fold_rows(rows, case["phi"])
    assert len(got) == len(case["expected"])
    for (tags, ts, v), (etag, ev) in zip(got, case["expected"]):
        assert tags == (etag,)
        ev = _fnum(ev)
        if np.isnan(ev):
            assert np.isnan(v)
        else:
            assert abs(v - ev) <= case["tol"] * max(abs(ev), 1.0) + (0.0 if case["tol"] else 0.0), (case["name"], v, ev)


def test_le_label_parse_follows_rust():
    lp = FOLD["le_parse"]
    for s, v in lp["finite"].items():
        assert orc.parse_f64_rust(s) == v, s
    for s in lp["inf"]:
        assert orc.parse_f64_rust(s) == float("inf"), s
    for s in lp["nan"]:
        assert np.isnan(orc.parse_f64_rust(s)), repr(s)


# ---------------------------------------------------------------------------------------------------------------
# tql/range.result: sum(rate()) / sum by (..)(rate()) on the reference's own table — the BASELINE config-3 query shape
# ---------------------------------------------------------------------------------------------------------------
import json as _json
import os as _os

SUM_RATE = _json.load(open(_os.path.join(_os.path.dirname(__file__), "golden", "reference_sum_rate_vectors.json")))


@pytest.mark.parametrize("case", SUM_RATE["cases"], ids=lambda c: c["name"])
def test_sum_rate_reference_tables(case):
    """rate() per series through the oracle's range path, then the by-label SUM (DataFusion's accumulator: plain f64 +=
    in series order), label matchers applied the way the scan would; every printed value of range.result reproduced to
    the last digit (they are the shortest round-trip decimals of the f64 results)."""
    keep = [s for s in SUM_RATE["series"] if all(s[k] == v for k, v in case["filter"].items())]
    if not keep:
        assert case["expected"] == []
        return
    ts = np.concatenate([np.array(s["ts"], np.int64) for s in keep])
    val = np.concatenate([np.array(s["val"], np.float64) for s in keep])
    offsets = np.concatenate([[0], np.cumsum([len(s["ts"]) for s in keep])]).astype(np.uint64)
    p = orc.make_params("rate", case["start"], case["end"], case["interval"], case["range"])
    out, valid = orc.range_query(p, ts, val, None, offsets, mode="faithful")
    keys = sorted({tuple(s[t] for t in case["by"]) for s in keep})
    gid = np.array([keys.index(tuple(s[t] for t in case["by"])) for s in keep], np.uint32)
    gsum, gcnt = orc.group_aggregate("sum", out, valid, gid, len(keys))
    T = orc.num_steps(case["start"], case["end"], case["interval"])
    got = []
    for g, key in enumerate(keys):
        for k in range(T):
            if gcnt[g, k]:
                got.append([dict(zip(case["by"], key)), case["start"] + k * case["interval"],
                            float(gsum[g, k]) * case.get("scale", 1.0)])
    assert got == case["expected"]


@pytest.mark.parametrize("case", SQL.get("function_cases", []), ids=lambda c: c["name"])
def test_sqlness_function_cases(case):
    """promql/functions.result (predict_linear, double_exponential_smoothing / holt_winters, quantile_over_time with scalar
    expressions as parameters) and the irate table of tql/operator.result (eval grid on half seconds, issue 5880): the
    printed rows, value for value."""
    names, ts, val, sid, offsets = pack_series(case["series"])
    p = orc.make_params(case["fn"], case["start"], case["end"], case["interval"], case["range"],
                        param0=case.get("param0", 0.0), param1=case.get("param1", 0.0))
    out, valid = orc.range_query(p, ts, val, sid, offsets, mode="faithful")
    T = orc.num_steps(case["start"], case["end"], case["interval"])
    vb = orc.valid_to_bool(valid, T)
    got = {(names[s], case["start"] + k * case["interval"]): float(out[s, k]) for s in range(len(names)) for k in range(T) if vb[s, k]}
    exp = {(n, t): fnum(v) for n, t, v in case["expected"]}
    assert set(got) == set(exp), (case["name"], sorted(got))
    for key, e in exp.items():
        assert got[key] == e, (case["name"], key, got[key], e)


def test_histogram_quantile_over_sum_by_le_across_partitions():
    """promql/histogram_multi_partition.result: histogram_quantile(0.5, sum by (le)(histogram_gap_bucket)) over a table
    partitioned by `shard` — instant selector per (shard, le) series, SUM by le merged across the two partitions
    (__sum_state / __sum_merge, commutativity.rs:85-113: partial sums and counts add), then the fold: 0.5 and
    0.5833333333333334.  The per-partition partials added together equal the single-pass aggregate."""
    rows = [(0, "0.5", "a", 1), (0, "1", "a", 2), (0, "+Inf", "a", 2), (0, "0.5", "z", 2), (0, "1", "z", 4), (0, "+Inf", "z", 4),
            (10000, "0.5", "a", 1), (10000, "1", "a", 2), (10000, "+Inf", "a", 2), (10000, "0.5", "z", 1), (10000, "1", "z", 3),
            (10000, "+Inf", "z", 3)]
    keys = sorted({(sh, le) for _, le, sh, _ in rows})                      # series = (shard, le), SeriesDivide order
    les = sorted({le for _, le, _, _ in rows}, key=lambda s: orc.parse_f64_rust(s))
    ts = np.array([t for k in keys for t, le, sh, v in rows if (sh, le) == k], np.int64)
    val = np.array([float(v) for k in keys for t, le, sh, v in rows if (sh, le) == k])
    offsets = np.arange(len(keys) + 1, dtype=np.uint64) * 2
    out, valid = orc.instant_query(ts, val, offsets, 0, 10000, 10000, 300000, 0)
    gid = np.array([les.index(le) for _, le in keys], np.uint32)
    gsum, gcnt = orc.group_aggregate("sum", out, valid, gid, len(les))
    # the two partitions (shard < 'n', shard >= 'n') aggregated on their own, then merged
    parts = []
    for pick in (lambda sh: sh < "n", lambda sh: sh >= "n"):
        idx = [i for i, (sh, _) in enumerate(keys) if pick(sh)]
        parts.append(orc.group_aggregate("sum", out[idx], valid[idx], gid[idx], len(les)))
    assert ((parts[0][0] + parts[1][0]) == gsum).all() and ((parts[0][1] + parts[1][1]) == gcnt).all()
    bounds = [orc.parse_f64_rust(le) for le in les]
    got = [orc.histogram_evaluate_row(0.5, bounds, gsum[:, k])[0] for k in range(2)]
    assert got == [0.5, 0.5833333333333334]


def test_tsid_regression_tables():
    """promql/tsid_histogram_quantile_regression.result: avg_over_time(m[5s]) at a 5 s step over samples at 0 / 5 / 10 s
    (each window (t - 5 s, t] holds exactly the sample at t: 1, 3, 5), and histogram_quantile(0.5, buckets 1 / 2 / +Inf)
    through the instant selector + fold: 1.5 at every step."""
    ts = np.array([0, 5000, 10000], np.int64)
    p = orc.make_params("avg_over_time", 0, 10000, 5000, 5000)
    out, valid = orc.range_query(p, ts, np.array([1.0, 3.0, 5.0]), None, np.array([0, 3], np.uint64), mode="faithful")
    assert orc.valid_to_bool(valid, 3).all() and out[0].tolist() == [1.0, 3.0, 5.0]
    # buckets '1', '2', '+Inf' of job1: counters 1,2,3 / 2,4,6 / 3,6,9
    bts = np.tile(ts, 3)
    bval = np.array([1, 2, 3, 2, 4, 6, 3, 6, 9], np.float64)
    sel, sv = orc.instant_query(bts, bval, np.array([0, 3, 6, 9], np.uint64), 0, 10000, 5000, 300000, 0)
    assert orc.valid_to_bool(sv, 3).all()
    bounds = [orc.parse_f64_rust(s) for s in ("1", "2", "+Inf")]
    assert [orc.histogram_evaluate_row(0.5, bounds, sel[:, k])[0] for k in range(3)] == [1.5, 1.5, 1.5]
