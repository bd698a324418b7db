#!/usr/bin/env bash
# 2-GPU call: TP-2 numerics for the three exchange routes (tests/test_tp_gpu.py) + the TP-2 bench line of the default route.
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/r02_tp2_final.sh'
set -u
mkdir -p gpurun_out
timeout 330 python -m pytest tests/test_tp_gpu.py -m gpu -x -q > gpurun_out/r02_tp2_tests.log 2>&1; echo "tp tests rc=$?"
grep -v "DEBUG\|torch.__trace" gpurun_out/r02_tp2_tests.log | tail -n 5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 64 --warmup 8 --no-cpu --no-extra > gpurun_out/r02_tp2_bench.json 2> gpurun_out/r02_tp2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    r = json.loads([l for l in open("gpurun_out/r02_tp2_bench.json") if l.startswith("{")][-1])
    print("tok/s", round(r["value"], 1), "e2e", round(r["e2e"]["value"], 1), r.get("tp_exchange"), "kernels/step", r.get("kernels_per_step"), "dense", r.get("dense_tok_s"))
except Exception as e:
    print("no result", e); print(open("gpurun_out/r02_tp2_bench.err").read()[-1500:])
PY
