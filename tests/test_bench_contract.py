"""bench.py's reference arm runs on the CPU (the unmodified reference naive path, oracle/_ref), so its JSON contract can be checked here:
exactly one line on stdout, the keys the driver reads, and the reference-arm conventions (impl, cpu_baseline, zero-byte e2e)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-sample", "1024"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "particle-steps/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["steps"] == 2 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the reference arm runs "on the measured arm's config" (bench contract): the same dictionary both arms print, from the command line alone;
    # what each step actually executed is said beside it
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    ours = bench.arm_config(argparse.Namespace(workload="plummer", n=1 << 24, order=4, leaf_capacity=48, scheme="auto", flags=0), 1)
    assert d["config"] == ours and d["config"]["workload"] == "plummer sphere N=16777216" and d["config"]["partition"] == "single"
    assert d["reference_sample"]["n"] == 1024 and "first 1024 particles" in d["cpu_baseline"]["sample"]
    assert bench.arm_config(argparse.Namespace(workload="plummer", n=1 << 24, order=4, leaf_capacity=48, scheme="auto", flags=0), 8)["flags"] == 128
    # BASELINE config 1 beside it: the unmodified NaiveSimulation on all 4096 particles of the uniform cube, 10 steps (the one
    # same-configuration ratio; our arm reports the same configuration under the same key)
    c1 = d["config1"]
    assert c1["same_config_as_ours_config1"] is True and c1["value"] > 0 and c1["unit"] == "particle-steps/s" and "4096" in c1["workload"]
    # ... and the reference's FMM kernels themselves, compiled for the host, on that configuration (informational)
    f = c1["reference_fmm_kernels_on_host"]
    assert ("unavailable" in f) or (f["value"] > 0 and f["cores"] == 1 and f["interaction_pairs"] == 327049)


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
