#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/quick_ab.py --steps 96 --repeats 3 --out gpurun_out/r02_quick_ab_resid0.jsonl 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gemv_gpu.py tests/test_decoder_gpu.py -x -q 2>&1 | tail -3
