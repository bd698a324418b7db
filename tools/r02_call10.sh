#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/bench_gemv.py --configs 0,60,61,64,65,66,67,68,69 --shapes 4096x14336,14336x4096,4096x4096,4096x6144 --out gpurun_out/r02_bench_gemv_gen5b.jsonl 2>&1 | grep -v dense_cublas | cut -c1-200
timeout 200 python tools/quick_ab.py --only configs --configs 60,61,64,65,66,67,68,69 --steps 96 --out gpurun_out/r02_quick_ab_gen5b.jsonl 2>&1 | tail -9
