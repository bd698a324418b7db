#!/usr/bin/env bash
# Round-2 first GPU call (trimmed from round2_first_call.sh): validate on hardware what round 1 wrote without a GPU, A/B it.
set -u
mkdir -p gpurun_out
export TEAL_SAMPLING_GPU_TESTS=1
timeout 150 python tools/quick_ab.py --only tunes --deadline 70 --steps 64 --out gpurun_out/r02_quick_ab.jsonl > /dev/null 2>&1; echo "quick_ab tunes rc=$?"
timeout 90 python tools/quick_ab.py --only chain --steps 64 --out gpurun_out/r02_quick_ab.jsonl > gpurun_out/r02_quick_ab_chain.log 2>&1; echo "quick_ab chain rc=$?"
cat gpurun_out/r02_quick_ab.jsonl
timeout 300 python -m pytest tests -m gpu -x -q -k "not gen2 and not unvalidated and not sampling_gpu" > gpurun_out/r02_gpu_tests_validated.log 2>&1; echo "validated suite rc=$?"
timeout 200 python -m pytest tests/test_sampling_gpu.py -q > gpurun_out/r02_gpu_tests_sampling.log 2>&1; echo "sampling tests rc=$?"
tail -n 5 gpurun_out/r02_gpu_tests_sampling.log
timeout 200 python tools/bench_gemv.py --configs 0 --shapes 4096x14336,14336x4096,4096x4096,4096x6144 --out gpurun_out/r02_bench_gemv.jsonl > /dev/null 2>&1; echo "bench_gemv rc=$?"
cat gpurun_out/r02_bench_gemv.jsonl
timeout 200 python tools/bench_ref_triton.py > gpurun_out/r02_ref_triton.log 2>&1; echo "ref triton rc=$?"; tail -n 6 gpurun_out/r02_ref_triton.log
timeout 200 python tools/rowchunk_timeline.py --per-launch --out gpurun_out/r02_rowchunk_timeline.jsonl > gpurun_out/r02_rowchunk_timeline.txt 2>&1; echo "timeline rc=$?"
tail -n 40 gpurun_out/r02_rowchunk_timeline.txt
timeout 200 python tools/rowchunk_timeline.py --chain mlp,attn,norms --out gpurun_out/r02_rowchunk_timeline_chained.jsonl > gpurun_out/r02_rowchunk_timeline_chained.txt 2>&1; echo "chained timeline rc=$?"
tail -n 40 gpurun_out/r02_rowchunk_timeline_chained.txt
