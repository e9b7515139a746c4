"""bench.py on the CPU: the algorithmic-work figures behind `roofline.achieved` and the JSON contract of the
reference arm (`--impl reference`: the oracle port on the host cores, bounded sample, the one place besides `cpu_baseline` where
bench.py executes oracle/).  No GPU is touched."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zig-tfhe_b200"))


def test_algorithmic_work_per_bootstrap_matches_the_survey():
    import bench
    import tfhe_b200
    P = tfhe_b200.PARAM_SETS
    # per 512-point transform 5*512*9 + 6*512 = 26,112 flop; per external product (2L+2) transforms + 2L*2*512 complex MACs x 8
    assert 5 * 512 * 9 + 6 * 512 == 26112
    flop, bsk, ksk = bench.work_per_bootstrap(P["128"])
    assert flop == 700 * 258048 == 180633600
    assert bsk == 68812800
    assert ksk == 103366656                      # with the never-read k = 0 rows, as the reference stores it
    assert bench.work_per_bootstrap(P["110"])[:2] == (162570240, 61931520)
    assert bench.work_per_bootstrap(P["80"])[:2] == (141926400, 54067200)
    assert bench.work_per_bootstrap(P["uint4"])[:2] == (820 * 120832, 26869760) == (99082240, 26869760)


def test_metric_names_follow_baseline_json():
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "gates" in bench.metric_name("128", "fast") and "128" in bench.metric_name("128", "fast")
    assert "UINT4" in bench.metric_name("uint4", "exact")
    assert bench.UNIT == "gates/s"
    assert isinstance(base.get("north_star", ""), str)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (what the driver runs next to the GPU arm): same metric / unit / config keys, impl = reference,
    e2e with zero copied bytes, cpu_baseline describing the run; a bounded sample, so it finishes in seconds"""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "gates/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("bootstrapped gates/sec") and d["value"] > 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"]
