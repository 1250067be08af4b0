"""CPU: the benchmark's reference arm (`bench.py --impl reference`: the unmodified reference package from oracle/_ref on
the host cores, or the oracle port where that copy is absent) runs without a GPU and prints ONE JSON line with the keys the
driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("no_ref", ["0", "1"])
def test_reference_arm_json_line(no_ref):
    have_ref = no_ref == "0" and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "octreelib"))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "120000"], capture_output=True, text=True, timeout=300, cwd=ROOT,
                         env=dict(os.environ, OL_NO_REFERENCE=no_ref))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("points/sec") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"] == "c4_street_100M"
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if have_ref else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "points" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
