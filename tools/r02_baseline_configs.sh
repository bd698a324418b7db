#!/usr/bin/env bash
# BASELINE.json configs 2, 3 (40 %), 5 on one GPU + the refreshed default line
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --steps 128 --warmup 16 --no-70b "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "$name rc=$?";
python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02_bench_$name.json") if l.startswith("{")][-1])
    print("  tok/s", round(r["value"],1), "dense", {k: round(v,1) for k,v in r["dense_tok_s"].items() if v}, "speedup", round(r["speedup_vs_dense"],3), "roofline frac", round(r["roofline"]["frac"],3), "kept", round(r["kept_fraction"],3), "sampled", r["sampled"] and round(r["sampled"].get("tok_s",0),1))
except Exception as e:
    print("  no result", e)
PY
}
run llama2_7b --workload llama-2-7b --no-cpu
run llama3_8b_s40 --workload llama-3-8b --sparsity 0.4 --no-cpu
run mistral_7b_greedy_means --workload mistral-7b --profile greedy-means --no-cpu
timeout 900 python bench.py --steps 128 --warmup 16 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "default rc=$?"
python - <<'PY'
import json
r = json.loads([l for l in open("gpurun_out/r02_bench.json") if l.startswith("{")][-1])
print({k: r[k] for k in ("value", "value_spread", "dense_tok_s", "speedup_vs_dense", "sampled", "tp_70b")})
print("roofline", {k: r["roofline"][k] for k in ("achieved", "frac", "us_per_launch", "traffic")}, "e2e", r["e2e"]["value"])
PY
