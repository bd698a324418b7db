#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_sampling_gpu.py tests/test_decoder_gpu.py -q 2>&1 | tail -3
timeout 200 python tools/gemv_timeline.py --out gpurun_out/r02_gemv_timeline.jsonl 2>&1 | tail -6
timeout 200 python tools/bench_gemv.py --configs 0 --shapes 4096x14336,14336x4096,4096x4096,4096x6144 --out gpurun_out/r02_bench_gemv_new.jsonl 2>&1 | cut -c1-230
