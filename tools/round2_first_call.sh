#!/usr/bin/env bash
# First GPU call of the next round: validate on hardware everything that was written after round 1's GPU budget was
# spent (it passes the CPU emulation of tests/emu, nothing more), then A/B it against the validated default.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Every step is bounded by its own timeout; results land in gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
export TEAL_PROGRAM_GPU_TESTS=1 TEAL_GEN2_GPU_TESTS=1 TEAL_SAMPLING_GPU_TESTS=1
# 0. the one-minute overview first (it survives a cut-off call): default vs flag-chained blocks / RMSNorm hand-offs vs
#    one-change-at-a-time variants vs generations 2 / 3, tokens/s + dense-mode logits check against the default kernels
timeout 150 python tools/quick_ab.py --only tunes --deadline 90 --steps 64 --out gpurun_out/r02_quick_ab.jsonl > /dev/null 2>&1; echo "quick_ab tunes rc=$?"
timeout 90 python tools/quick_ab.py --only chain --steps 64 --out gpurun_out/r02_quick_ab.jsonl > /dev/null 2>&1; echo "quick_ab chain rc=$?"
cat gpurun_out/r02_quick_ab.jsonl
# 1. validated suite first (must stay green), then the opt-in tests of the unvalidated kernels
timeout 600 python -m pytest tests -m gpu -x -q -k "not gen2 and not unvalidated and not sampling_gpu" > gpurun_out/r02_gpu_tests_validated.log 2>&1; echo "validated suite rc=$?"
timeout 900 python -m pytest tests/test_gemv_gen2_gpu.py -q > gpurun_out/r02_gpu_tests_gen2.log 2>&1; echo "gen2/gen3 tests rc=$?"
timeout 300 python -m pytest tests/test_sampling_gpu.py -q > gpurun_out/r02_gpu_tests_sampling.log 2>&1; echo "sampling tests rc=$?"
timeout 300 python -m pytest tests/test_step_program_gpu.py -q -k unvalidated > gpurun_out/r02_gpu_tests_program_variants.log 2>&1; echo "step-program variants rc=$?"
# 2. decode-step A/B on Llama-3-8B @50 %: graph engine under every dispatch setting, then the step-program sweeps
timeout 420 python tools/bench_engines.py --quick --graph-tunes --out gpurun_out/r02_graph_tunes.jsonl > /dev/null 2>&1; echo "graph tunes rc=$?"
timeout 300 python tools/bench_engines.py --sweep 3 --out gpurun_out/r02_program_sweep3.jsonl > /dev/null 2>&1; echo "sweep3 rc=$?"
timeout 300 python tools/bench_engines.py --sweep 4 --out gpurun_out/r02_program_sweep4.jsonl > /dev/null 2>&1; echo "sweep4 rc=$?"
# 3. per-projection micro-benchmark: validated variants vs generation 2 / 3
timeout 300 python tools/bench_gemv.py --configs 31,32,42,43,44,45,46,50,51,52 --shapes 4096x14336,14336x4096,4096x4096,4096x6144 \
    --out gpurun_out/r02_bench_gemv.jsonl > /dev/null 2>&1; echo "bench_gemv rc=$?"
# 4. where the step's time goes: in-kernel timeline of the production kernels inside the graph replay
timeout 240 python tools/rowchunk_timeline.py --per-launch --out gpurun_out/r02_rowchunk_timeline.jsonl > gpurun_out/r02_rowchunk_timeline.txt 2>&1; echo "timeline rc=$?"
timeout 240 python tools/rowchunk_timeline.py --chain mlp,attn,norms --out gpurun_out/r02_rowchunk_timeline_chained.jsonl > gpurun_out/r02_rowchunk_timeline_chained.txt 2>&1; echo "chained timeline rc=$?"
# 5. is FHFMA (generation 2's arithmetic) slower than conversions + FFMA2?
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fhfma_rate tools/micro/fhfma_rate.cu && timeout 60 /tmp/fhfma_rate > gpurun_out/r02_micro_fhfma_rate.txt 2>&1; echo "fhfma_rate rc=$?"
tail -n 40 gpurun_out/r02_graph_tunes.jsonl
# Multi-GPU (separate call, N x the box time): the exchange fused into the GEMV finalisers vs the one-shot kernel
#   gpurun --gpus 2 --timeout 600 -- 'for m in fused gemv; do TEAL_TP_ALLREDUCE=$m python -m torch.distributed.run --nnodes=1 \
#     --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 --no-cpu --no-dense \
#     > gpurun_out/r02_tp2_$m.json; done; python tools/tp_check.py'
