#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/bench_gemv.py --configs 0,54,55,56,57,58,59 --shapes 4096x14336,14336x4096,4096x4096,4096x6144 --out gpurun_out/r02_bench_gemv_ring.jsonl 2>&1 | grep -v dense_cublas | cut -c1-220
timeout 120 python tools/quick_ab.py --only tunes --deadline 20 --steps 64 --out gpurun_out/r02_quick_ab_ring.jsonl 2>&1 | tail -8
