#!/usr/bin/env bash
# usage: tools/r02_tp8.sh N : TP-N numerics (fused-in-GEMV exchange), then Llama-3-8B and Llama-2-70B bench lines
set -u
mkdir -p gpurun_out
N=${1:-8}
TEAL_TP_ALLREDUCE=gemv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/tp_check.py --out gpurun_out/r02_tp${N}_check.jsonl > gpurun_out/r02_tp${N}_check_gemv.log 2>&1; echo "tp_check rc=$?"
grep -E "^\{|TP_CHECK" gpurun_out/r02_tp${N}_check_gemv.log | cut -c1-230 | head -8
for wl in llama-3-8b llama-2-70b; do
  echo "== bench N=$N $wl"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 64 --warmup 8 --no-cpu --no-extra --workload $wl > gpurun_out/r02_tp${N}_bench_$wl.json 2> gpurun_out/r02_tp${N}_bench_$wl.err; echo "rc=$?"
  python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02_tp${N}_bench_$wl.json") if l.startswith("{")][-1])
    print("tok/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), r["config"].get("tp_exchange"), "kernels/step", r.get("kernels_per_step"), "dense", r.get("dense_tok_s"), "spread", r.get("value_spread"))
except Exception as e:
    print("no result", e); print(open("gpurun_out/r02_tp${N}_bench_$wl.err").read()[-1500:])
PY
done
