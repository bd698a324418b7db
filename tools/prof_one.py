"""Tiny driver for ncu: a few launches of one projection shape."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import teal_b200.ops as ops

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="4096x14336")
ap.add_argument("--configs", default="0")
ap.add_argument("--sparsity", type=float, default=0.5)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--gate-up", action="store_true")
a = ap.parse_args()
for shp in a.shapes.split(","):
    Z, N = (int(v) for v in shp.split("x"))
    Ws = [(torch.randn(Z, N, device="cuda", dtype=torch.float16) * 0.02).T for _ in range(a.reps + 1)]
    x = torch.randn(1, 1, Z, device="cuda", dtype=torch.float16)
    thr = float(x.abs().float().flatten().kthvalue(max(1, int(Z * a.sparsity))).values) if a.sparsity > 0 else -1.0
    for cfg in (int(c) for c in a.configs.split(",")):
        ops.set_config(cfg)
        for i in range(a.reps):
            if a.gate_up:
                ops.sparse_gate_up(x, Ws[i], Ws[i + 1], thr, thr)
            else:
                ops.splitk_sparse_gemv(x, Ws[i], thr, 0)
        torch.cuda.synchronize()
