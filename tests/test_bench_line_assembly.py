"""bench.py's measured arm needs a B200; what can be checked without one is the bookkeeping around the measurement: argument
handling, the stage-time arithmetic and the assembly of the ONE JSON line with every key the driver and the judge read.
This test runs bench.main() in a subprocess in which the device library is replaced by a stub that only RECORDS calls and
returns fixed stage times (tests/host/bench_stub_runner.py) — nothing is measured and nothing here is a fallback of the
product (the product has none: tests/test_abi.py::test_no_cpu_fallback). It exists so that an edit to bench.py cannot
break the round-end run unnoticed on a box without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_measured_arm_assembles_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host", "bench_stub_runner.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == base["metric"] and d["unit"] == "particle-steps/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    for key in ("value", "ms_per_step", "device_ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e",
                "gpu_launches", "roofline", "other_rooflines", "cpu_baseline", "accuracy", "config1", "reference_capacity", "stage_ms", "counts"):
        assert key in d, key
    assert d["scaling"] == "strong" and d["dtype"] == "f32" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"] == "plummer sphere N=65536" and d["config"]["leaf_capacity"] == 48 and "l2_policy" in d["config"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"):
        assert key in d["roofline"], key
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] == 48 * 65536
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["gpu_launches"] == 2 * 121                      # 121 launches per step for a 12-level tree (profiles/r02j launch list)
    assert d["other_rooflines"]["keys_sort_gather"]["bound"] == "hbm" and d["other_rooflines"]["m2l"]["bound"] == "fp32"
    # the same dictionary the reference arm prints for the same command line
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    assert d["config"] == bench.arm_config(argparse.Namespace(workload="plummer", n=65536, order=4, leaf_capacity=48, scheme="auto", flags=0), 1)
