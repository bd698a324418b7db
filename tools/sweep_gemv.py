"""The reference's kernel micro-benchmark (scripts/benchmark_gemv.py:186-258) on B200: one projection over sparsity
0 ... 0.95, ours vs the reference's Triton kernel (oracle/_ref, when fetched) vs dense cuBLAS vs "(1 - s) x dense"
(the reference's "theoretical optimal" line, :218-223).  Inputs as in the reference: x ~ U(-0.5, 0.5), threshold s/2
(:166,199-200).  CUDA-graph chains over rotating weight copies (L2 defeated), CUDA events.

    python tools/sweep_gemv.py [--shapes 4096x14336,14336x4096] [--step 0.05] [--out gpurun_out/sweep_gemv.jsonl]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import teal_b200.ops as ops  # noqa: E402
from tools.bench_gemv import PEAK, time_graph  # noqa: E402

ref_gemv = None
REF = os.path.join(ROOT, "oracle", "_ref")
if os.path.isdir(os.path.join(REF, "kernels")):
    sys.path.insert(0, REF)
    try:
        from kernels.sparse_gemv import splitk_sparse_gemv as ref_gemv  # the reference's Triton kernel, unmodified
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"ref_triton": f"import failed: {e}"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4096x14336,14336x4096")
    ap.add_argument("--step", type=float, default=0.05)
    ap.add_argument("--out", default="gpurun_out/sweep_gemv.jsonl")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")
    for shp in args.shapes.split(","):
        Z, N = (int(v) for v in shp.split("x"))
        ncopies = max(2, min(32, (768 << 20) // (Z * N * 2)))
        Ws = [(torch.randn(Z, N, device="cuda", dtype=torch.float16) * 0.02).T for _ in range(ncopies)]
        x = (torch.rand(1, 1, Z, device="cuda") - 0.5).half()
        dense_us, _ = time_graph([(lambda i=i: torch.matmul(x, Ws[i % ncopies].T)) for i in range(32)])
        s = 0.0
        while s < 0.951:
            thr = s / 2
            nnz = int((x.abs() > torch.tensor(thr, dtype=torch.float16, device="cuda")).sum())
            alg = nnz * N * 2 + Z * 2 + N * 2
            ours, _ = time_graph([(lambda i=i: ops.splitk_sparse_gemv(x, Ws[i % ncopies], thr, 0)) for i in range(32)])
            rec = dict(shape=shp, sparsity=round(s, 2), nnz=nnz, ours_us=round(ours, 2), dense_cublas_us=round(dense_us, 2),
                       theoretical_optimal_us=round((1 - s) * dense_us, 2), ours_gbs=round(alg / ours / 1e3),
                       ours_frac_of_measured_peak=round(alg / ours / 1e3 / PEAK, 3))
            if ref_gemv is not None:
                try:
                    for _ in range(2):
                        ref_gemv(x, Ws[0], thr, int(s // 0.05))  # autotune warm-up (key includes the sparsity bin)
                    ref_us, _ = time_graph([(lambda i=i: ref_gemv(x, Ws[i % ncopies], thr, int(s // 0.05))) for i in range(32)])
                    rec["ref_triton_us"] = round(ref_us, 2)
                    rec["speedup_vs_ref_triton"] = round(ref_us / ours, 2)
                except Exception as e:  # noqa: BLE001
                    rec["ref_triton"] = f"failed: {type(e).__name__}"
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")
            s += args.step
        del Ws
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
