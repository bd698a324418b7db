#!/usr/bin/env bash
# 2-GPU call: TP numerics on hardware (TP-2 logits == single-GPU logits of the same state dict) for the three exchange
# implementations, then tok/s of each.
set -u
mkdir -p gpurun_out
N=${1:-2}
for m in fused nccl gemv; do
  echo "== tp_check N=$N exchange=$m"
  TEAL_TP_ALLREDUCE=$m timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/tp_check.py > gpurun_out/r02_tp${N}_check_$m.log 2>&1; echo "rc=$?"
  grep -E "max\|tp|tokens|TP_CHECK|fused peer|Error|error" gpurun_out/r02_tp${N}_check_$m.log | head -30
done
for m in fused gemv nccl; do
  echo "== bench N=$N exchange=$m"
  TEAL_TP_ALLREDUCE=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 64 --warmup 8 --no-cpu --no-dense > gpurun_out/r02_tp${N}_bench_$m.json 2> gpurun_out/r02_tp${N}_bench_$m.err; echo "rc=$?"
  python - <<PY
import json
try:
    r=json.loads([l for l in open("gpurun_out/r02_tp${N}_bench_$m.json") if l.startswith("{")][-1])
    print("tok/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), r["config"].get("tp_exchange"), "kernels/step", r.get("kernels_per_step"))
except Exception as e:
    print("no result", e)
PY
done
