#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ref_caller_gpu.py tests/test_decoder_gpu.py tests/test_ref_triton_gpu.py -x -q > gpurun_out/r02_gpu_tests3.log 2>&1; echo "new gpu tests rc=$?"; tail -n 15 gpurun_out/r02_gpu_tests3.log | cut -c1-300
timeout 300 python tools/bench_ref_triton.py > gpurun_out/r02_ref_triton.log 2>&1; echo "ref triton rc=$?"; tail -n 5 gpurun_out/r02_ref_triton.log | cut -c1-300
timeout 300 python tools/fixture_run.py --out gpurun_out/r02_fixture_run.jsonl 2>&1 | tail -2 | cut -c1-700
timeout 600 python tools/sweep_gemv.py --out gpurun_out/r02_sweep_gemv.jsonl > gpurun_out/r02_sweep.log 2>&1; echo "sweep rc=$?"; tail -n 4 gpurun_out/r02_sweep.log | cut -c1-300
