#!/usr/bin/env bash
# usage: tools/r02_70b.sh N : Llama-2-70B bench line at N GPUs (N = 1: plain python)
set -u
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu --no-extra --workload llama-2-70b > gpurun_out/r02_tp1_bench_llama-2-70b.json 2> gpurun_out/r02_tp1_bench_llama-2-70b.err; echo "rc=$?"
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 64 --warmup 8 --no-cpu --no-extra --workload llama-2-70b > gpurun_out/r02_tp${N}_bench_llama-2-70b.json 2> gpurun_out/r02_tp${N}_bench_llama-2-70b.err; echo "rc=$?"
fi
python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02_tp${N}_bench_llama-2-70b.json") if l.startswith("{")][-1])
    print("tok/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), r.get("tp_exchange"), "dense", r.get("dense_tok_s"), "spread", r.get("value_spread"), "roofline", r.get("roofline") and {k: r["roofline"][k] for k in ("frac","us_per_launch")}, "step_hbm_frac", r.get("step_hbm_frac"))
except Exception as e:
    print("no result", e); print(open("gpurun_out/r02_tp${N}_bench_llama-2-70b.err").read()[-1500:])
PY
