"""The bench.py JSON contract, checked on the committed lines of the last GPU round (profiles/) and on a live
run of the CPU reference arm (a tiny sample)."""
import glob
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _last(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_committed_bench_line_has_every_contract_key():
    d = _last("r02_v*_bench.json")
    assert BASE_KEYS <= set(d) and {"roofline", "cpu_baseline", "clocks"} <= set(d)
    assert d["metric"].startswith("folded nt/sec") and d["unit"] == "nt/s" and d["higher_is_better"] is True
    assert d["scaling"] == "strong" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"] and "configs[2]" in d["config"]["workload"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert abs(d["value"] - d["config"]["nt"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    assert d["parity_in_run"]["equal"] is True and d["parity_in_run"]["text_bytes"] == 1235240762
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("reference", "port")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_multi_gpu_lines_are_strong_scaling_through_one_context():
    one = _last("r02_v*_bench.json")
    for n in (2, 4, 8):
        d = _last("r02_v*_bench_%dgpu.json" % n)
        assert d["n_gpus"] == n and d["scaling"] == "strong" and d["config"]["devices"] == n
        assert d["config"]["nt"] == one["config"]["nt"]                       # the same job on more GPUs
        assert d["parity_in_run"]["equal"] is True                           # the multi-device result is the RNALfold text
        assert d["sharding"]["lpt_imbalance"] < 1.001
        assert d["e2e"]["value"] < d["value"] and d["weak"]["scaling"] == "weak"
        assert d["value"] > 0.85 * n * one["value"]                          # resident strong-scaling efficiency


def test_committed_reference_line():
    d = _last("r02_v*_bench_reference.json")
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["gpu_launches"] == 0


def test_reference_arm_runs_here():
    """bench.py --impl reference on a tiny sample (the reference's RNALfold, or the oracle port without it)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-loci-per-core", "1"], stdout=subprocess.PIPE, check=True, timeout=600).stdout.decode()
    d = json.loads(out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "nt/s" and d["steps"] == 1
