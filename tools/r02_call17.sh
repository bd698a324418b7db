#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_decoder_gpu.py -x -q 2>&1 | tail -3
for pf in 1 0; do
  echo "== TEAL_PREFETCH_O=$pf"
  TEAL_PREFETCH_O=$pf timeout 300 python tools/quick_ab.py --steps 96 --repeats 3 --out gpurun_out/r02_quick_ab_attn_pf$pf.jsonl 2>&1 | tail -1
done
