#!/usr/bin/env bash
# One GPU call = what the driver does at round end: parity suite (-m gpu), smoke(), then the default bench line.
#   gpurun --timeout 1800 -- 'bash tools/r02_gpu_suite.sh'
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; grep -v "DEBUG\|torch.__trace" gpurun_out/r02_gpu_tests.log | tail -n 4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02_bench.err
python - <<'PY'
import json
try:
    r = json.loads([l for l in open("gpurun_out/r02_bench.json") if l.startswith("{")][-1])
    print({k: r[k] for k in ("value", "ms_per_step", "value_spread", "dense_tok_s", "speedup_vs_dense", "sampled", "tp_70b", "kernels_per_step", "step_hbm_frac")})
    print("e2e", r["e2e"], "roofline", {k: r["roofline"][k] for k in ("achieved", "frac", "us_per_launch", "kept_fraction", "traffic")})
    print("cpu", r["cpu_baseline"], "clocks", r["clocks"])
except Exception as e:
    print("no bench line", e)
PY
