#!/usr/bin/env bash
# usage: tools/r02_tp_n.sh N : TP-N numerics record (exchange fused into the GEMV) + the default bench line (incl. the 70B leg) at N GPUs
set -u
mkdir -p gpurun_out
N=${1:-8}
TEAL_TP_ALLREDUCE=gemv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/tp_check.py --out gpurun_out/r02_tp${N}_check.jsonl > gpurun_out/r02_tp${N}_check_gemv.log 2>&1; echo "tp_check rc=$?"
grep -E "^\{|TP_CHECK" gpurun_out/r02_tp${N}_check_gemv.log | cut -c1-260 | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 64 --warmup 8 --no-cpu > gpurun_out/r02_tp${N}_bench_default.json 2> gpurun_out/r02_tp${N}_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02_tp${N}_bench_default.json") if l.startswith("{")][-1])
    print("tok/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), r.get("tp_exchange"), "dense", r.get("dense_tok_s"), "sampled", r.get("sampled"), "tp_70b", r.get("tp_70b"))
except Exception as e:
    print("no result", e); print(open("gpurun_out/r02_tp${N}_bench_default.err").read()[-1500:])
PY
