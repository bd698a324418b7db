#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; grep -v "DEBUG\|torch.__trace" gpurun_out/r02_gpu_tests.log | tail -n 6
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
