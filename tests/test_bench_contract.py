"""bench.py's reference arm (`--impl reference`: the reference's CPU implementation of the path on the host cores) prints
the one JSON line the driver parses, with the keys of the bench contract, and finishes quickly on a bounded sample.
The GPU arm is exercised by the driver itself (and by tools/r02_gpu_suite.sh); this runs on the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")     # what torchrun sets: the arm must undo it and use the host cores
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["higher_is_better"] is True and rec["unit"] == "tok/s"
    assert rec["metric"].startswith("decode tok/s batch=1 llama-3-8b")
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in rec, k
    assert rec["value"] > 0 and abs(rec["ms_per_step"] * rec["value"] - 1e3) < 1.0          # whole tokens, wall time / steps
    assert "workload" in rec["config"] and "model" not in rec["config"]
    cb = rec["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == rec["value"] and cb["sample"]
    assert cb["cores"] == (os.cpu_count() or 1) or cb["cores"] > 1 or (os.cpu_count() or 1) == 1
    assert rec["e2e"] == {"value": rec["value"], "unit": rec["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
