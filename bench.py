#!/usr/bin/env python
"""bench.py — PromQL rate() range-query throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path  (one JSON line on rank 0)
  python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path (oracle port)

A "step" is one pass of the hot path (K0 series offsets + the fused normalize/range/rate stage: K2L first tier,
K2 over the series it hands off, slow kernel over what K2 hands off) over
one HBM-resident chunk of synthetic series of the BASELINE config-2 shape: 1000 samples/series at a
15 s scrape (+<1 s jitter), rate(x[5m]) at a 15 s step => 1000 eval steps.  Config 2's 10 M series
(200 GB of input) exceed one GPU's HBM, so they are processed as 8 chunks of 1.25 M series; the
default K = 8 timed steps are exactly one 10 M-series job.  `value` is input samples/s with inputs
resident in HBM; `e2e` is the same metric through the host-pointer C-ABI call (pinned host buffers,
H2D + kernels + D2H inside the timed region).  Inputs (25 GB) are far larger than L2 (126 MB), so
no explicit L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T0 = 1_700_000_000_000
N_SAMPLES = 1000
SCRAPE = 15_000
RANGE = int(os.environ.get("B2P_BENCH_RANGE_MS", "300000"))  # 5m lookback (override: tuning experiments only)
SEED = 0x5EED
METRIC = "rate() input samples/sec"
UNIT = "samples/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    The sampler is started before the warm-up (nvidia-smi takes a few hundred ms to produce its first line) and
    every line carries a timestamp; stop(t0, t1) keeps the samples that fall inside the timed window [t0, t1]
    (wall clock), or — if the window was shorter than the sampling period — the samples nearest to it."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = []
        for seen, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                import datetime
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                ts = seen
            try:
                rows.append((ts, float(f[2]), float(f[3]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        where = "inside the timed region"
        if not inside and rows and t0 is not None:
            mid = 0.5 * (t0 + t1)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            where = "nearest to the timed region (region shorter than the sampling period)"
        if t0 is None:
            inside = rows
        sm = [r[1] for r in inside]
        mx = [r[2] for r in inside]
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where}


def query_params():
    from greptimedb_b200 import make_params
    return make_params("rate", T0, T0 + (N_SAMPLES - 1) * SCRAPE, SCRAPE, RANGE)


def cpu_reference_pass(n_series: int, threads: int, faithful: bool = True, series_begin: int = 0, with_resets=0):
    """One pass of the reference's CPU algorithm (oracle port, structure-faithful) -> (seconds, samples)."""
    import numpy as np
    from oracle import oracle as orc
    ts, val, sid = orc.synth_fill(series_begin, n_series, N_SAMPLES, T0, SCRAPE, 1000, with_resets, SEED)
    offsets = np.arange(n_series + 1, dtype=np.uint64) * N_SAMPLES
    p = orc.make_params("rate", T0, T0 + (N_SAMPLES - 1) * SCRAPE, SCRAPE, RANGE)
    t = time.perf_counter()
    orc.range_query(p, ts, val, sid, offsets, mode="faithful" if faithful else "flat", threads=threads)
    return time.perf_counter() - t, n_series * N_SAMPLES


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust/DataFusion build is
    impossible here (no rustc/cargo/network), so this times the oracle port — the C restatement of
    SeriesDivide -> SeriesNormalize -> RangeManipulate -> prom_rate -> Filter, structure-faithful
    (per-series batch materialisation, packed RangeArray keys, tag take, null filter) — on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores * 1024, 4096)
    per_step = min(per_step, 262_144)
    for _ in range(args.warmup):
        cpu_reference_pass(per_step, cores)
    t = 0.0
    samples = 0
    for _ in range(args.steps):
        dt, n = cpu_reference_pass(per_step, cores)
        t += dt
        samples += n
    value = samples / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"rate(x[5m]) step 15s over {per_step} series x {N_SAMPLES} samples per step "
                               "(bounded sample of BASELINE config 2), CPU oracle port of the reference path",
                   "series_per_step": per_step, "samples_per_series": N_SAMPLES},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x {per_step} series x {N_SAMPLES} samples, structure-faithful port"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from greptimedb_b200 import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S = args.series_per_gpu
    n_rows = S * N_SAMPLES
    T = N_SAMPLES
    Tw = (T + 31) // 32
    p = query_params()

    ctx = Context(local)
    ctx.use_torch_stream()
    ts = torch.empty(n_rows, dtype=torch.int64, device=dev)
    val = torch.empty(n_rows, dtype=torch.float64, device=dev)
    sid = torch.empty(n_rows, dtype=torch.int32, device=dev)
    offsets = torch.empty(S + 1, dtype=torch.int64, device=dev)
    out = torch.empty(S * T, dtype=torch.float64, device=dev)
    valid = torch.empty(S * Tw, dtype=torch.int32, device=dev)
    # series are hash-sharded across GPUs: rank r owns global series [r*S, (r+1)*S) of this step's chunk
    ctx.synth_fill_dev(rank * S, S, N_SAMPLES, T0, SCRAPE, 1000, args.resets, SEED, ts, val, sid)
    ctx.sync()
    torch.cuda.synchronize()

    sumby = args.workload == "sumby"
    if sumby:
        # BASELINE config 3: sum by (pod)(rate(x[5m])), 100 k label groups over the whole 10 M-series job;
        # every rank reduces its shard into [G x T] (sum, cnt) partials, ONE all-reduce per buffer merges them
        from greptimedb_b200 import distributed as D
        G = args.groups
        gid_np = (D.mix32(np.arange(rank * S, (rank + 1) * S, dtype=np.uint32)) % np.uint32(G)).astype(np.int32)
        gid = torch.from_numpy(gid_np).to(dev)
        gsum = torch.zeros(G * T, dtype=torch.float64, device=dev)
        gcnt = torch.zeros(G * T, dtype=torch.int32, device=dev)

    def step():
        ctx.series_offsets_dev(sid, n_rows, S, offsets)
        ctx.range_eval_dev(p, ts, val, offsets, n_rows, S, out, valid)
        if sumby:
            ctx.group_aggregate_dev("sum", out, valid, gid, S, G, T, gsum, gcnt)
            if world > 1:
                dist.all_reduce(gsum)
                dist.all_reduce(gcnt)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    ctx.sync()
    slow_series = ctx.last_slow_series()
    warp_tier_series = ctx.last_warp_tier_series()
    barrier()
    launches0 = ctx.launch_count()
    wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0_ms, k2_ms = [], []
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    ctx.sync()
    warp_tier_series = ctx.last_warp_tier_series()   # series the first tier handed to K2 in the last timed step
    barrier()
    wall1 = time.time()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    # per-kernel durations: re-run K untimed-by-the-headline steps reading the library's own CUDA events
    for _ in range(min(args.steps, 5)):
        step()
        ctx.sync()
        k0_ms.append(ctx.kernel_ms(0))
        k2_ms.append(ctx.kernel_ms(1) )
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    samples_per_step_all = S * N_SAMPLES * world
    value = samples_per_step_all * args.steps / (elapsed_ms / 1e3)

    # ---- end to end through the host-pointer C ABI: pinned host buffers, H2D + kernels + D2H timed ----
    Se = args.e2e_series
    e2e = None
    if Se > 0:
        h_ts = torch.empty(Se * N_SAMPLES, dtype=torch.int64).pin_memory()
        h_val = torch.empty(Se * N_SAMPLES, dtype=torch.float64).pin_memory()
        h_sid = torch.empty(Se * N_SAMPLES, dtype=torch.int32).pin_memory()
        h_out = torch.empty(Se * T, dtype=torch.float64).pin_memory()
        h_valid = torch.empty(Se * Tw, dtype=torch.int32).pin_memory()
        h_ts.copy_(ts[: Se * N_SAMPLES])
        h_val.copy_(val[: Se * N_SAMPLES])
        h_sid.copy_(sid[: Se * N_SAMPLES])
        torch.cuda.synchronize()
        import ctypes as C
        L = ctx._L

        def e2e_step():
            rc = L.b2p_range_eval(ctx._h, C.byref(p), C.c_void_p(h_ts.data_ptr()), C.c_void_p(h_val.data_ptr()),
                                  C.c_void_p(h_sid.data_ptr()), None, Se * N_SAMPLES, Se,
                                  C.c_void_p(h_out.data_ptr()), C.c_void_p(h_valid.data_ptr()), None)
            if rc != 0:
                raise RuntimeError(L.b2p_last_error().decode())

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        n_e2e = max(2, min(args.steps, 4))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()          # synchronous: returns after the D2H of the result
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": Se * N_SAMPLES * world * n_e2e / dt, "unit": UNIT,
               "h2d_bytes_per_step": Se * N_SAMPLES * 20, "d2h_bytes_per_step": Se * T * 8 + Se * Tw * 4,
               "series_per_step": Se, "steps": n_e2e}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    k2 = statistics.mean(k2_ms) if k2_ms else float("nan")
    k0 = statistics.mean(k0_ms) if k0_ms else float("nan")
    alg_k2 = 16.0 * n_rows + 8.0 * S * T + 4.0 * S * Tw + 8.0 * (S + 1)   # bytes per launch of the dominant kernel
    achieved = alg_k2 / (k2 * 1e-3) / 1e9
    lean_on = os.environ.get("B2P_DISABLE_LEAN_TIER", "0") != "1" and os.environ.get("B2P_ENABLE_THREAD_TIER", "0") != "1"
    if lean_on and warp_tier_series * 2 > S:
        # the first tier still declined most series in the timed steps: K2 did the work
        lean_on = False
        kernel_name = "range_fast_kernel<rate> (the lean first tier handed off most series)"
    elif lean_on and args.resets:
        kernel_name = ("range_lean_kernel<rate, bit words> (adaptive: the plain variant handed off every series during the "
                       "warm-up) + range_fast_kernel<rate> over the series it hands off")
    else:
        kernel_name = ("range_lean_kernel<rate> (+ range_fast_kernel<rate> over the series it hands off)" if lean_on
                       else "range_fast_kernel<rate>")
    step_ms = elapsed_ms / args.steps
    read_frac = 20.0 * n_rows / (step_ms * 1e-3) / 1e9 / peak
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": (f"sum by(pod)(rate(x[5m])) -> {args.groups} groups + all-reduce, " if sumby else "") +
                               f"rate(x[5m]) step 15s over {S} series x {N_SAMPLES} samples per GPU per step "
                               f"(BASELINE config 2 = 10M series processed as chunks of {S}); resets={args.resets}",
                   "series_per_gpu_per_step": S, "samples_per_series": N_SAMPLES, "eval_steps": T,
                   "parallelism": f"series-sharded x{world}, " + ("one all-reduce of [G x T] (sum f64, cnt i32) per step"
                                                                  if sumby else "no data-path collective"),
                   "l2": "inputs (16-25 GB per step) >> 126 MB L2; no flush needed"},
        "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
                     # captures (2e8-sample launch), scaled: K2L 3.214 + 1.599 GB (profiles/r1_range_lean_kernel.md),
                     # K2 alone 3.381 + 1.585 GB (profiles/r1_range_fast_kernel.md, version f)
                     "traffic": (4.813e9 if lean_on else 4.966e9) * (n_rows / 2.0e8), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_k2, "kernel_ms": k2, "k0_series_offsets_ms": k0,
                     "hbm_read_frac_whole_step": read_frac},
        "gpu_launches": launches, "slow_path_series": slow_series, "warp_tier_series": warp_tier_series, "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = min(max(cores * 2048, 8192), 262_144)
        dt, n = cpu_reference_pass(n_cpu, cores, faithful=True)
        dt_flat, n_flat = cpu_reference_pass(n_cpu, cores, faithful=False)
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_cpu} series x {N_SAMPLES} samples, structure-faithful oracle port, "
                                          f"{cores} threads", "algorithm_only_value": n_flat / dt_flat}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--series-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--e2e-series", type=int, default=131_072)
    ap.add_argument("--resets", type=int, default=0, help="1 = counter-reset variant of the value generator")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="rate", choices=["rate", "sumby"],
                    help="rate = BASELINE config 2 (headline); sumby = config 3: + by-label sum and one all-reduce")
    ap.add_argument("--groups", type=int, default=100_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
