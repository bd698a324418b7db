#!/usr/bin/env bash
# ncu evidence for the committed kernels (1 GPU): launch list of the timed sparse step, then --set full of one layer's four GEMVs
set -u
mkdir -p gpurun_out
TEAL_NCU_WINDOW=1 timeout 170 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02_launches_decode_step_ncu.csv python bench.py --steps 2 --warmup 3 --no-dense --no-cpu --no-extra --repeats 1 > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
TEAL_NCU_WINDOW=1 timeout 150 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sparse_gemv_kernel -s 40 -c 4 \
  -o gpurun_out/r02_prof_gemv_layer -f python bench.py --steps 2 --warmup 3 --no-dense --no-cpu --no-extra --repeats 1 > gpurun_out/r02_ncu_full.log 2>&1; echo "set full rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_decode_step_ncu.csv
