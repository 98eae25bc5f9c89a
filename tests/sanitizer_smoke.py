"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel once, on irregular
series, both ring variants, the slow path, the lean first tier on regular series (pair groups, early finish,
hand-off) and the opt-in thread tier.  Not collected by pytest (no test_ prefix);
run on a GPU box:  compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from greptimedb_b200 import Context, make_params  # noqa: E402
from tests.test_gpu_parity import QUERY_SHAPES, make_irregular  # noqa: E402


def main():
    ts, val, offsets = make_irregular(1234, 48)
    for env in ("0", "1"):
        os.environ["B2P_ENABLE_THREAD_TIER"] = env
        ctx = Context(0)
        for fn in ("rate", "resets", "sum_over_time", "quantile_over_time", "deriv", "absent_over_time"):
            for q in QUERY_SHAPES:
                p = make_params(fn, q["start"], q["end"], q["interval"], q["range"], offset=q["offset"], param0=0.5)
                ctx.range_eval(p, ts, val, offsets=offsets)
        day = 86_400_000
        p = make_params("rate", 1_000_000, 1_000_000 + 40 * day, 3_600_000, 2 * day)   # int64 ring
        ctx.range_eval(p, ts, val, offsets=offsets)
        out, valid = ctx.instant_select(ts, val, 1_000_000, 4_000_000, 15_000, 300_000, 0, offsets=offsets)
        S, T = out.shape
        gid = (np.arange(S) % 5).astype(np.uint32)
        gs, gc = ctx.group_aggregate("sum", out, valid, gid, 5)
        B = 4
        ctx.histogram_quantile(0.9, np.array([0.1, 1.0, 5.0, np.inf]), out[: (S // B) * B], valid[: (S // B) * B])
        ctx.close()
    # regular series of the BASELINE shape: the lean tier's steady state (pair groups), its early finish (query ends
    # before the data), history before the query, and the hand-off of series with resets / NaN samples
    from oracle import oracle as orc
    S, N, T0 = 64, 1000, 1_700_000_000_000
    ctx = Context(0)
    for resets in (0, 1):
        ts, val, sid = orc.synth_fill(0, S, N, T0, 15_000, 1000, resets, 0x5EED)
        val = val.copy()
        val.reshape(S, N)[::8, 700] = np.nan
        for fn in ("rate", "delta", "avg_over_time", "last_over_time"):
            for start, end, interval, rng in ((T0, T0 + 999 * 15_000, 15_000, 300_000),
                                              (T0 + 3_000_000, T0 + 6_000_000, 5_000, 60_000),
                                              (T0 - 600_000, T0 + 999 * 15_000 + 900_000, 45_000, 300_000)):
                ctx.range_eval_n(make_params(fn, start, end, interval, rng), ts, val, sid, None, S)
    ctx.close()
    print("sanitizer smoke done")


if __name__ == "__main__":
    main()
