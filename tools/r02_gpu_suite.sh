#!/usr/bin/env bash
# One GPU call = what the driver does at round end: the default bench line, smoke(), the parity suite (-m gpu).
#   gpurun --timeout 1800 -- 'bash tools/r02_gpu_suite.sh'
# (bench first: if the call is cut short by the GPU budget the bench line is already on disk)
set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02_bench.err
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
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout ${TEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
grep -v "DEBUG\|torch.__trace" gpurun_out/r02_gpu_tests.log | tail -n 20
