"""CPU property test of the arithmetic claim behind the first tier's uniform-cadence variant
(greptimedb_b200/csrc/b2p_kernel_lean.cuh, lean_pair / lean_value_shape): on a series sampled exactly at the eval
interval every uncut window has the same shape, ExtrapolatedRate::calc's value-independent tail
(extrapolate_rate.rs:246-284) evaluated ONCE for that shape is a constant `factor`, and

    rate = result_value * factor        bit for bit

whenever calc leaves duration_to_start alone (no counter, result <= 0, first < 0, or — the kernel's exact shortcut —
first >= result and sampled >= duration_to_start, which makes duration_to_zero >= duration_to_start).  Checked here in
IEEE f64 (numpy) against the oracle's restatement of calc over random regular series, phases, ranges and steps of the
grid; the device-side twin is tests/test_gpu_parity.py::test_uniform_cadence_tier_matches_oracle_and_the_general_tiers."""
import numpy as np
import pytest

from oracle import oracle as orc


def shape_factor(fn, length, to_start, to_end, cadence, range_ms):
    """extrapolate_rate.rs:246-284 with duration_to_start untouched, in the reference's operation order."""
    sampled = np.float64((length - 1) * cadence)
    average = sampled / np.float64(length - 1)
    threshold = average * np.float64(1.1)
    ext = sampled
    ext = ext + (np.float64(to_start) if np.float64(to_start) < threshold else average / np.float64(2.0))
    ext = ext + (np.float64(to_end) if np.float64(to_end) < threshold else average / np.float64(2.0))
    factor = ext / sampled
    if fn == "rate":
        factor = factor / (np.float64(range_ms) / np.float64(1000.0))
    return factor


@pytest.mark.parametrize("fn", ["rate", "increase", "delta"])
def test_result_times_shape_factor_equals_calc_bit_for_bit(fn):
    rng = np.random.default_rng(20260921)
    checked = shortcut = 0
    for trial in range(60):
        cadence = int(rng.choice([1000, 15_000, 30_000, 7_001]))
        n = int(rng.integers(40, 400))
        t0 = 1_700_000_000_000 + int(rng.integers(0, cadence))
        range_ms = int(rng.choice([2, 5, 20, 61])) * cadence + int(rng.integers(0, cadence))
        start = t0 + int(rng.integers(-3, 30)) * cadence + int(rng.integers(0, cadence))   # any phase against the samples
        T = int(rng.integers(5, 200))
        ts = (t0 + np.arange(n) * cadence).astype(np.int64)
        # counters with small and large bases (both sides of first >= result), deltas on signed walks
        if fn == "delta":
            val = np.cumsum(rng.normal(size=n) * 5.0)
        else:
            val = np.cumsum(rng.random(n) * rng.choice([0.5, 50.0])) + rng.choice([0.0, 3.0, 1e4])
        p = orc.make_params(fn, start, start + (T - 1) * cadence, cadence, range_ms)
        out, valid = orc.range_query(p, ts, val, None, np.array([0, n], np.uint64), mode="faithful")
        vb = orc.valid_to_bool(valid, T)[0]
        for k in range(T):
            te = start + k * cadence
            tlo = te - range_ms
            inside = np.flatnonzero((ts > tlo) & (ts <= te))
            if inside.size < 2:
                continue
            lo, hi = int(inside[0]), int(inside[-1])
            if lo == 0 or hi == n - 1:
                continue                      # windows cut by an end of the series take calc itself in the kernel
            assert vb[k]
            first, last = val[lo], val[hi]
            result = last - first
            to_start, to_end = int(ts[lo] - tlo), int(te - ts[hi])
            sampled_i = int(ts[hi] - ts[lo])
            if fn != "delta" and result > 0.0 and first >= 0.0:
                if not (first >= result and sampled_i >= to_start):
                    continue                  # the zero crossing may move duration_to_start: calc itself
                shortcut += 1
                # the shortcut's premise, checked in exact arithmetic on the f64 operands
                assert np.float64(sampled_i) * (first / result) >= np.float64(to_start)
            f = shape_factor(fn, hi - lo + 1, to_start, to_end, cadence, range_ms)
            assert (np.float64(result) * f).tobytes() == np.float64(out[0, k]).tobytes(), (fn, trial, k)
            checked += 1
    assert checked > 2000 and (fn == "delta" or shortcut > 500), (checked, shortcut)


def test_all_uncut_windows_of_a_regular_series_share_one_shape():
    """Translation invariance: with samples exactly one eval interval apart, window k+1 is window k moved on by one sample."""
    rng = np.random.default_rng(7)
    for _ in range(50):
        cadence = int(rng.choice([1000, 15_000, 60_000]))
        n = 300
        t0 = int(rng.integers(0, 10**9))
        ts = t0 + np.arange(n) * cadence
        range_ms = int(rng.integers(cadence, 30 * cadence))
        start = t0 + 40 * cadence + int(rng.integers(0, cadence))
        shapes = set()
        for k in range(200):
            te = start + k * cadence
            inside = np.flatnonzero((ts > te - range_ms) & (ts <= te))
            shapes.add((inside.size, int(te - ts[inside[-1]]), int(ts[inside[0]] - (te - range_ms))))
        assert len(shapes) == 1
