"""CPU checks of bench.py's contract pieces that do not need a GPU: the committed ncu traffic file carries the kernel
variants the JSON line's roofline.traffic is read from, and the reference arm (--impl reference: the oracle port timed on
the host cores) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_traffic_file_has_the_kernel_variants_the_bench_reads():
    sys.path.insert(0, ROOT)
    import bench
    for key in ("range_lean_kernel", "range_lean_kernel_uniform"):
        per_sample, src = bench.load_traffic(key)
        assert per_sample is not None and src, key
        # ts 8 + val 8 read, 8 B + 1 bit written per step (steps ~ samples in config 2): ~24 B per input sample
        assert 20.0 < per_sample < 30.0, (key, per_sample)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"]
