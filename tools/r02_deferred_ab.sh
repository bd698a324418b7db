#!/usr/bin/env bash
# Hardware A/B of the deferred epilogue (QKV / gate+up leave fp32 sums, their consumers round them) against the route
# where every launch finalises its own tiles; then the parity tests that cover it.
#   gpurun --timeout 900 -- 'bash tools/r02_deferred_ab.sh'
set -u
mkdir -p gpurun_out
rm -f gpurun_out/quick_ab_deferred.jsonl
timeout 400 python tools/quick_ab.py --ab-deferred 3 --repeats 2 --steps 128 --out gpurun_out/quick_ab_deferred.jsonl 2> gpurun_out/quick_ab_deferred.err | tail -n 8
tail -n 3 gpurun_out/quick_ab_deferred.err
timeout 500 python -m pytest tests/test_decoder_gpu.py tests/test_gemv_gpu.py tests/test_ref_caller_gpu.py -m gpu -x -q > gpurun_out/r02_deferred_tests.log 2>&1; echo "tests rc=$?"
grep -v "DEBUG\|torch.__trace" gpurun_out/r02_deferred_tests.log | tail -n 6
