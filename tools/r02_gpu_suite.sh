#!/usr/bin/env bash
# one GPU call: parity suite, then the bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 6 gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py --steps 128 --warmup 16 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02_bench.err
python - <<'PY'
import json
try:
    r = json.loads([l for l in open("gpurun_out/r02_bench.json") if l.startswith("{")][-1])
    print({k: r[k] for k in ("value", "ms_per_step", "value_spread", "dense_tok_s", "speedup_vs_dense", "sampled", "kernels_per_step", "step_hbm_frac")})
    print("e2e", r["e2e"], "roofline", {k: r["roofline"][k] for k in ("achieved", "frac", "us_per_launch", "kept_fraction")})
    print("cpu", r["cpu_baseline"])
except Exception as e:
    print("no bench line", e)
PY
