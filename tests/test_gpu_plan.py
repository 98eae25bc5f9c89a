"""GPU: the C++ plan layer (GpuPromRangeExec) fed with Arrow RecordBatches, written like the reference's own
operator tests (in-memory batch -> exec node -> collected rows; range_manipulate.rs:839-907,
promql_test.rs:343-405), checked against the sqlness goldens and the oracle."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as orc
from tests.helpers import fnum, load_sqlness

pytestmark = pytest.mark.gpu
SQL = load_sqlness()


@pytest.fixture(scope="module")
def ctx():
    from greptimedb_b200 import Context
    c = Context(0)
    yield c
    c.close()


def batch_from_series(series, tag="x", split=None):
    ts, val, tags = [], [], []
    for name, d in series.items():
        ts += d["ts"]
        val += [fnum(v) for v in d["val"]]
        tags += [name] * len(d["ts"])
    b = pa.record_batch([pa.array(ts, pa.timestamp("ms")), pa.array(val, pa.float64()), pa.array(tags, pa.string())],
                        names=["ts", "val", tag])
    if split:
        return [b.slice(0, split), b.slice(split)]
    return [b]


def test_offset_rate_sqlness_through_the_plan(ctx):
    """tql eval (3000,3000,'1s') rate(calculate_rate_offset_total[10m] offset 5m)  — promql/offset.result:104-112"""
    from greptimedb_b200.plan import PromRangeExec
    case = next(c for c in SQL["range_cases"] if c["name"] == "offset_rate_10m_offset_5m")
    for split in (None, 7, 11):   # series boundaries inside and across input batches
        ex = PromRangeExec(ctx, "prom_rate", case["start"], case["end"], case["interval"], case["range"], "ts", "val",
                           ["x"], offset=case["offset"])
        for b in batch_from_series(case["series"], split=split):
            ex.push(b)
        out = ex.execute()
        assert ex.num_series() == 2
        assert out.schema.names == ["ts", "prom_rate(ts_range,val)", "x"]
        rows = list(zip(out.column(2).to_pylist(), [int(t.timestamp() * 1000) for t in out.column(0).to_pylist()],
                        out.column(1).to_pylist()))
        assert rows == [tuple(r) for r in case["expected"]]


def test_nan_series_and_null_filter_rows(ctx):
    """min_over_time over the NaN fixture: `only_nan` emits no row at all (tql/aggr_over_time.result)."""
    from greptimedb_b200.plan import PromRangeExec
    case = next(c for c in SQL["range_cases"] if c["name"] == "min_over_time_nan")
    ex = PromRangeExec(ctx, "prom_min_over_time", case["start"], case["end"], case["interval"], case["range"], "ts",
                       "val", ["ty"])
    for b in batch_from_series(SQL["series_sets"]["nan_data"], tag="ty"):
        ex.push(b)
    out = ex.execute()
    got = sorted(zip(out.column(2).to_pylist(), out.column(1).to_pylist()))
    assert got == sorted((n, v) for n, _, v in case["expected"])
    assert ex.num_series() == 5


def test_series_divide_reference_fixture_through_the_plan(ctx):
    """series_divide.rs:668-905: three RecordBatches, two Utf8 tag columns (multi-byte values), one series spanning all
    three batches -> 7 series with the reference's row counts (count_over_time over everything, one eval step)."""
    from greptimedb_b200.plan import PromRangeExec
    from tests.helpers import load_unit
    g = load_unit()["series_divide"]
    ex = PromRangeExec(ctx, "prom_count_over_time", 23_000, 23_000, 1_000, 60_000, "ts", "val", g["tag_columns"])
    for b in g["batches"]:
        n = len(b["ts"])
        ex.push(pa.record_batch([pa.array(b["ts"], pa.timestamp("ms")), pa.array([1.0] * n, pa.float64()),
                                 pa.array(b["host"], pa.string()), pa.array(b["path"], pa.string())],
                                names=["ts", "val", "host", "path"]))
    out = ex.execute()
    assert ex.num_series() == 7
    got = list(zip(out.column("host").to_pylist(), out.column("path").to_pylist(), out.column(1).to_pylist()))
    assert got == [(e["host"], e["path"], float(e["rows"])) for e in g["expected_series"]]


def test_sum_by_plan_matches_oracle_and_is_sorted(ctx):
    """sum by (pod)(rate(http_requests_total[5m])) on synthetic series: rows sorted by (label, ts)."""
    from greptimedb_b200.plan import PromRangeExec
    S, N, T0 = 60, 120, 1_700_000_000_000
    ts, val, sid = orc.synth_fill(0, S, N, T0, 15_000, 1000, 1, 7)
    pods = [f"pod-{s % 7}" for s in range(S)]
    insts = [f"i{s:03d}" for s in range(S)]
    order = sorted(range(S), key=lambda s: (pods[s], insts[s]))        # scan order: sorted by the tag tuple
    rows = np.concatenate([np.arange(s * N, (s + 1) * N) for s in order])
    b = pa.record_batch([pa.array(ts[rows], pa.timestamp("ms")), pa.array(val[rows]),
                         pa.array(np.repeat([pods[s] for s in order], N)),
                         pa.array(np.repeat([insts[s] for s in order], N))], names=["ts", "v", "pod", "instance"])
    ex = PromRangeExec(ctx, "prom_rate", T0, T0 + (N - 1) * 15_000, 30_000, 300_000, "ts", "v", ["pod", "instance"],
                       aggregate="sum", by_columns=["pod"])
    ex.push(b.slice(0, 1000))
    ex.push(b.slice(1000))
    out = ex.execute()
    assert out.schema.names == ["pod", "ts", "sum(prom_rate)"]
    got_keys = list(zip(out.column(0).to_pylist(), [int(t.timestamp() * 1000) for t in out.column(1).to_pylist()]))
    assert got_keys == sorted(got_keys)
    # oracle: same rows, same series order
    offsets = np.arange(S + 1, dtype=np.uint64) * N
    p = orc.make_params("rate", T0, T0 + (N - 1) * 15_000, 30_000, 300_000)
    e_out, e_valid = orc.range_query(p, ts[rows], val[rows], None, offsets)
    names = sorted(set(pods))
    gid = np.array([names.index(pods[s]) for s in order], np.uint32)
    e_sum, e_cnt = orc.group_aggregate("sum", e_out, e_valid, gid, len(names))
    exp = {(names[g], T0 + k * 30_000): e_sum[g, k] for g in range(len(names)) for k in range(e_sum.shape[1]) if e_cnt[g, k]}
    got = dict(zip(got_keys, out.column(2).to_pylist()))
    assert set(got) == set(exp)
    for k in exp:
        assert abs(got[k] - exp[k]) <= 1e-9 * abs(exp[k])


def test_tsid_key_and_errors(ctx):
    from greptimedb_b200 import B2PError
    from greptimedb_b200.plan import PromRangeExec
    b = pa.record_batch([pa.array([0, 10_000, 20_000, 0, 10_000], pa.timestamp("ms")), pa.array([1.0, 2.0, 4.0, 5.0, 5.0]),
                         pa.array([7, 7, 7, 9, 9], pa.uint64())], names=["ts", "greptime_value", "__tsid"])
    ex = PromRangeExec(ctx, "prom_delta", 20_000, 20_000, 1000, 60_000, "ts", "greptime_value", ["__tsid"])
    ex.push(b)
    out = ex.execute()
    assert out.column(2).to_pylist() == [7, 9] and out.column(1).to_pylist() == [3.75, 0.0]
    with pytest.raises(B2PError):
        PromRangeExec(ctx, "prom_nope", 0, 1, 1, 1, "ts", "v", [])
    ex2 = PromRangeExec(ctx, "prom_rate", 0, 1, 1, 1, "timestamp", "v", [])
    with pytest.raises(B2PError) as ei:
        ex2.push(b)
    assert "No field named timestamp" in str(ei.value)


def test_histogram_quantile_through_the_plan(ctx):
    """tql eval (2820,2820,'1s') histogram_quantile(phi, rate(histogram2_bucket[15m])) — simple_histogram.result:237-262:
    SeriesDivide by the `le` tag, rate per bucket series, HistogramFold on top."""
    from greptimedb_b200.plan import PromRangeExec
    case = next(c for c in SQL["histogram_cases"] if c["name"] == "histogram2_rate_15m")
    les = ["0", "2", "4", "6", "+Inf"]
    # the scan delivers rows sorted by the primary key (le as a STRING: "+Inf" < "0" < "2" ...) then ts
    order = sorted(range(len(les)), key=lambda i: les[i])
    ts, val, le = [], [], []
    for i in order:
        ts += case["bucket_ts"]
        val += case["bucket_val"][i]
        le += [les[i]] * len(case["bucket_ts"])
    b = pa.record_batch([pa.array(ts, pa.timestamp("ms")), pa.array(val, pa.float64()), pa.array(le)],
                        names=["ts", "val", "le"])
    for q, expected in case["quantiles"]:
        ex = PromRangeExec(ctx, "prom_rate", case["start"], case["end"], case["interval"], case["range"], "ts", "val",
                           ["le"], histogram_quantile=q)
        ex.push(b)
        out = ex.execute()
        assert out.schema.names == ["ts", "prom_rate(ts_range,val)"]
        assert out.column(1).to_pylist() == [expected], (q, out.column(1).to_pylist())


def test_instant_selector_through_the_plan(ctx):
    """tql eval (3000,3000,'1s') calculate_rate_offset_total offset 10m — promql/offset.result:85-92."""
    from greptimedb_b200.plan import PromRangeExec
    case = next(c for c in SQL["instant_cases"] if c["name"] == "instant_offset_10m_at_3000")
    ex = PromRangeExec(ctx, "", case["start"], case["end"], case["interval"], 0, "ts", "val", ["x"],
                       offset=case["offset"], lookback_delta=case["lookback"])
    for bb in batch_from_series(case["series"], split=5):
        ex.push(bb)
    out = ex.execute()
    assert out.schema.names == ["ts", "val", "x"]
    rows = list(zip(out.column(2).to_pylist(), [int(t.timestamp() * 1000) for t in out.column(0).to_pylist()],
                    out.column(1).to_pylist()))
    assert rows == [tuple(r) for r in case["expected"]]


# ---------------------------------------------------------------------------------------------------
# by-label aggregators: the reference's result tables (tests-integration/src/tests/promql_test.rs:343-665) through the
# plan layer — instant selector (lookback 0) + Aggregate + Sort on the GPU
# ---------------------------------------------------------------------------------------------------
def _aggr_golden():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_aggregator_vectors.json")) as f:
        return json.load(f)


AGGR = _aggr_golden()


@pytest.mark.parametrize("case", AGGR["cases"], ids=lambda c: c["name"])
def test_by_label_aggregators_reference_tables_through_the_plan(ctx, case):
    from greptimedb_b200.plan import PromRangeExec
    sel = [s for s in AGGR["series"] if all(s[k] == v for k, v in case["filter"].items())]   # the label matcher
    sel.sort(key=lambda s: tuple(s[t] for t in AGGR["tags"]))                                   # scan order
    cols = {t: [] for t in AGGR["tags"]}
    ts, val = [], []
    for s in sel:
        ts += s["ts"]
        val += s["val"]
        for t in AGGR["tags"]:
            cols[t] += [s[t]] * len(s["ts"])
    b = pa.record_batch([pa.array(ts, pa.timestamp("ms")), pa.array(val, pa.float64())] + [pa.array(cols[t]) for t in AGGR["tags"]],
                        names=[AGGR["time_index"], AGGR["field"]] + AGGR["tags"])
    ex = PromRangeExec(ctx, "", AGGR["start"], AGGR["end"], AGGR["interval"], 0, AGGR["time_index"], AGGR["field"],
                       AGGR["tags"], lookback_delta=AGGR["lookback"], aggregate=case["agg"], by_columns=case["by"])
    ex.push(b)
    out = ex.execute()
    nby = len(case["by"])
    assert out.schema.names[:nby] == case["by"]
    keys = list(zip(*[out.column(i).to_pylist() for i in range(nby)])) if nby else [()] * out.num_rows
    tss = [int(t.timestamp() * 1000) for t in out.column(nby).to_pylist()]
    got = {(tuple(k), t): v for k, t, v in zip(keys, tss, out.column(nby + 1).to_pylist())}
    assert list(zip(keys, tss)) == sorted(zip(keys, tss))        # .sort(group_exprs asc) — planner.rs:443-449
    exp = {(tuple(lbl[bcol] for bcol in case["by"]), int(t)): float(v) for lbl, t, v in case["expected"]}
    assert set(got) == set(exp), (sorted(got), sorted(exp))
    tol = case.get("rel_tol", 0.0)
    for k in exp:
        assert abs(got[k] - exp[k]) <= tol * abs(exp[k]), (k, got[k], exp[k])


def test_histogram_fold_plan_mixed_layouts_and_strict_le_parse(ctx):
    """Histograms with different bucket layouts in one query (the reference's safe mode, histogram_fold.rs:834-846) and le
    labels only Rust's f64 parse accepts: ' 2' (leading blank) and '0x4' are NaN bounds -> the row is NaN; '+inf', '1e0'
    parse.  Buckets arrive in the scan's string order and are folded in numeric le order."""
    from greptimedb_b200.plan import PromRangeExec
    T0, N = 1_700_000_000_000, 40
    tsv = [T0 + i * 15_000 for i in range(N)]
    hists = {"a": ["0.5", "1e0", "2", "+inf"], "b": ["1", "+Inf"], "c": ["0.5", " 2", "+Inf"], "d": ["0.5", "1"]}
    ts, val, host, le = [], [], [], []
    for hname, les in sorted(hists.items()):
        for j, l in enumerate(sorted(les)):                      # primary-key (string) order within the histogram
            rank = sorted(les, key=lambda x: (np.isnan(orc.parse_f64_rust(x)), orc.parse_f64_rust(x))).index(l)
            for i in range(N):
                ts.append(tsv[i])
                val.append(float((rank + 1) * (i + 1) * (1 + ord(hname) % 3)))
                host.append(hname)
                le.append(l)
    b = pa.record_batch([pa.array(ts, pa.timestamp("ms")), pa.array(val, pa.float64()), pa.array(host), pa.array(le)],
                        names=["ts", "val", "host", "le"])
    start, end, step, rng = T0 + 300_000, T0 + (N - 1) * 15_000, 60_000, 300_000
    ex = PromRangeExec(ctx, "prom_rate", start, end, step, rng, "ts", "val", ["host", "le"], histogram_quantile=0.5)
    ex.push(b)
    out = ex.execute()
    assert out.schema.names == ["ts", "prom_rate(ts_range,val)", "host"]
    got = {}
    for t, v, hname in zip(out.column(0).to_pylist(), out.column(1).to_pylist(), out.column(2).to_pylist()):
        got[(hname, int(t.timestamp() * 1000))] = v
    # expectation: rate per bucket series from the oracle, then the row-literal fold
    rows = []
    for hname, les in sorted(hists.items()):
        for l in les:
            sel = [i for i in range(len(ts)) if host[i] == hname and le[i] == l]
            st, sv = np.array([ts[i] for i in sel], np.int64), np.array([val[i] for i in sel])
            op = orc.make_params("rate", start, end, step, rng)
            o, v = orc.range_query(op, st, sv, None, np.array([0, st.size], np.uint64))
            for k in range(o.shape[1]):
                if (v[0, k >> 5] >> (k & 31)) & 1:
                    rows.append(((hname,), start + k * step, l, o[0, k]))
    rows.sort(key=lambda r: (r[0], r[1], np.isnan(orc.parse_f64_rust(r[2])), orc.parse_f64_rust(r[2]) if not np.isnan(orc.parse_f64_rust(r[2])) else 0.0))
    exp = {(t[0], tsx): v for t, tsx, v in orc.histogram_fold_rows(rows, 0.5)}
    assert set(got) == set(exp)
    for k in exp:
        e, g = exp[k], got[k]
        assert (np.isnan(e) and np.isnan(g)) or abs(e - g) <= 1e-9 * abs(e), (k, e, g)
    assert all(np.isnan(v) for (hname, _), v in got.items() if hname in ("c", "d"))      # NaN bound / no +Inf bucket
    assert not any(np.isnan(v) for (hname, _), v in got.items() if hname in ("a", "b"))
