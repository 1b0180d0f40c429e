"""`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): runs without a GPU, prints ONE JSON line with the
contract's keys, on the same metric / unit / config.workload as the product arm, with the cpu_baseline description of itself and an
e2e object that repeats its own value (no device, no copies)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-rows", "300000", "--segment-rows", "300000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "decoded+filtered rows/sec" and d["unit"] == "rows/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("cfg3")
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
