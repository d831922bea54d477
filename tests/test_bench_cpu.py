"""bench.py without a GPU: the reference arm (the C++ / OpenMP port of the oracle, oracle/cpu_port.cpp) prints the
contract's JSON line for the driver's command line, honours --steps / --warmup, and only rank 0 works under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_line_c1():
    out = _run(["--impl", "reference", "--workload", "c1", "--gpus", "1", "--steps", "4", "--warmup", "1"])
    j = json.loads(out.splitlines()[-1])
    assert j["impl"] == "reference" and j["unit"] == "iterations/s" and j["higher_is_better"] is True
    assert j["steps"] == 4 and j["warmup"] == 1 and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port-c++" and j["cpu_baseline"]["cores"] >= 1
    assert "full workload" in j["cpu_baseline"]["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["vs_baseline"] is None and j["config"]["workload"].startswith("c1")


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--workload", "c1", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out == ""


def test_workload_table_matches_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfgs) == 5
    w = bench.WORKLOADS
    assert (w["c1"]["K"], w["c1"]["G"], w["c1"]["N"]) == (96, 100, 5) and "96×100" in cfgs[0].replace("\\u00d7", "×")
    assert (w["c2"]["G"], w["c2"]["N"], w["c2"]["learn"], w["c2"]["MH"]) == (500, 10, True, True)
    assert (w["c3"]["G"], w["c3"]["N"], w["c3"]["sharded"]) == (100000, 20, True)
    assert (w["c4"]["likelihood"], w["c4"]["G"], w["c4"]["N"]) == ("normal", 20000, 15)
    assert (w["c5"]["K"], w["c5"]["G"], w["c5"]["N"], w["c5"]["prior"]) == (1536, 50000, 40, "exponential")
    assert bench.metric_name(w["c3"]) == "Gibbs iterations/s (Poisson-Gamma, K=96, G=100000, N=20)"
    # algorithmic bytes of the north-star kernel at C3, fp64 state: the figure DESIGN.md quotes
    assert bench.algorithmic_bytes("k_zstat", w["c3"], 100000, 8) == 62430720
