#!/usr/bin/env bash
# N-GPU call: TP numerics on hardware (TP-N logits == single-GPU logits of the same model) for the three exchange
# implementations, then tok/s of each.   usage: tools/r02_tp.sh N [workloads...]
set -u
mkdir -p gpurun_out
N=${1:-2}
for m in gemv fused nccl; do
  echo "== tp_check N=$N exchange=$m"
  TEAL_TP_ALLREDUCE=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/tp_check.py --out gpurun_out/r02_tp${N}_check.jsonl > gpurun_out/r02_tp${N}_check_$m.log 2>&1; echo "rc=$?"
  grep -E "^\{|TP_CHECK|Error|error" gpurun_out/r02_tp${N}_check_$m.log | cut -c1-260 | head -12
done
for m in gemv fused nccl; do
  echo "== bench N=$N exchange=$m"
  TEAL_TP_ALLREDUCE=$m timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 64 --warmup 8 --no-cpu --no-extra > gpurun_out/r02_tp${N}_bench_$m.json 2> gpurun_out/r02_tp${N}_bench_$m.err; echo "rc=$?"
  python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02_tp${N}_bench_$m.json") if l.startswith("{")][-1])
    print("tok/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), r.get("tp_exchange"), "kernels/step", r.get("kernels_per_step"), "dense", r.get("dense_tok_s"))
except Exception as e:
    print("no result", e); print(open("gpurun_out/r02_tp${N}_bench_$m.err").read()[-1500:])
PY
done
