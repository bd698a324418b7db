#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for pct in 100 90 80 70 60 50 120 150 200; do
  echo "== TEAL_WAVE_PCT=$pct"
  TEAL_WAVE_PCT=$pct timeout 120 python tools/quick_ab.py --steps 96 --repeats 2 --out gpurun_out/r02_quick_ab_wave$pct.jsonl 2>&1 | tail -1
done
