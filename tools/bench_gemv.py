"""Per-projection micro-benchmark of the fused sparse GEMV (run on the GPU box).

Methodology follows the reference's scripts/benchmark_gemv.py + kernels/kernel_utils.py (CUDA-event
timing, L2 defeated) but L2 is defeated by rotating over distinct weight copies totalling >= 384 MB
and launches are replayed from a CUDA graph so the host is never the bottleneck.
Algorithmic bytes per call: nnz*N*2 + Z*2 + N*2.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import teal_b200.ops as ops  # noqa: E402

PEAK = 6572.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def time_graph(fn_list, iters=5):
    """fn_list: callables to capture back-to-back in one graph; returns us per callable."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fn_list[:2]:
            f()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for f in fn_list:
                f()
        g.replay()
        torch.cuda.synchronize()
        best = 1e30
        tot = 0.0
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay()
            e1.record(s)
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e3 / len(fn_list)
            best = min(best, t)
            tot += t
    return tot / iters, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4096x14336,14336x4096,4096x4096,4096x6144,4096x11008,8192x28672")
    ap.add_argument("--configs", default="0", help="teal_set_config ids: (variant + 1) + 16 * (lanes per warp / 32), 0 = by size")
    ap.add_argument("--sparsities", default="0.5")
    ap.add_argument("--out", default="gpurun_out/bench_gemv.jsonl")
    ap.add_argument("--launches", type=int, default=48)
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    dev = torch.device("cuda")
    out = open(args.out, "a")
    for shp in args.shapes.split(","):
        Z, N = (int(v) for v in shp.split("x"))
        per = Z * N * 2
        ncopies = max(2, min(48, (768 << 20) // per))
        Ws = [(torch.randn(Z, N, device=dev, dtype=torch.float16) * 0.02).T for _ in range(ncopies)]
        x = torch.randn(1, 1, Z, device=dev, dtype=torch.float16)
        for sp in (float(s) for s in args.sparsities.split(",")):
            if sp <= 0:
                thr = -1.0
            else:
                thr = float(x.abs().float().flatten().kthvalue(max(1, int(Z * sp))).values)
            nnz = int((x.abs() > torch.tensor(thr, dtype=torch.float16, device=dev)).sum()) if sp > 0 else Z
            alg = nnz * N * 2 + Z * 2 + N * 2
            for cfg in (int(c) for c in args.configs.split(",")):
                ops.set_config(cfg)
                fns = [(lambda i=i: ops.splitk_sparse_gemv(x, Ws[i % ncopies], thr, 0)) for i in range(args.launches)]
                avg, best = time_graph(fns)
                rec = dict(kind="sparse_gemv", Z=Z, N=N, sparsity=sp, nnz=nnz, cfg=cfg, us_avg=round(avg, 3),
                           us_best=round(best, 3), alg_bytes=alg, gbs=round(alg / avg / 1e3, 1),
                           frac_of_measured_peak=round(alg / avg / 1e3 / PEAK, 4), copies=ncopies)
                print(json.dumps(rec), flush=True)
                out.write(json.dumps(rec) + "\n")
            ops.set_config(0)
        # dense cuBLAS comparator on the same layout (x @ W^T with W column-major == x @ Wt)
        fns = [(lambda i=i: torch.matmul(x, Ws[i % ncopies].T)) for i in range(args.launches)]
        avg, best = time_graph(fns)
        alg = Z * N * 2 + Z * 2 + N * 2
        rec = dict(kind="dense_cublas", Z=Z, N=N, us_avg=round(avg, 3), us_best=round(best, 3), alg_bytes=alg,
                   gbs=round(alg / avg / 1e3, 1), frac_of_measured_peak=round(alg / avg / 1e3 / PEAK, 4))
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
        del Ws
        torch.cuda.empty_cache()
    out.close()


if __name__ == "__main__":
    main()
