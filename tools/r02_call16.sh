#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/quick_ab.py --configs 2,3 --steps 96 --repeats 2 --out gpurun_out/r02_quick_ab_ssqrep.jsonl 2>&1 | tail -4
timeout 200 python tools/gemv_timeline.py --out gpurun_out/r02_gemv_timeline2.jsonl 2>&1 | tail -6 | cut -c1-900
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests2.log 2>&1; echo "gpu tests rc=$?"; tail -n 4 gpurun_out/r02_gpu_tests2.log
