#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${1:-8}
TEAL_TP_ALLREDUCE=gemv timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  tools/tp_check.py --out gpurun_out/r02_tp${N}_check.jsonl > gpurun_out/r02_tp${N}_check_gemv.log 2>&1; echo "tp_check rc=$?"
grep -E "^\{|TP_CHECK" gpurun_out/r02_tp${N}_check_gemv.log | cut -c1-260 | head -8
grep -E "Error" gpurun_out/r02_tp${N}_check_gemv.log | head -3
