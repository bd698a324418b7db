"""In-kernel phase timeline of the fused GEMV (uses teal_set_debug_timestamps)."""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import teal_b200.ops as ops
from teal_b200 import _lib

NAMES = ["start", "x_staged", "scan", "pieces", "lists", "first_full", "consumed", "red_bar", "counter",
         "epilogue", "end", "prod_first", "prod_done"]
ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="4096x4096,4096x14336")
ap.add_argument("--configs", default="0")
ap.add_argument("--sparsity", type=float, default=0.5)
a = ap.parse_args()
lib = _lib.load()
nsm = lib.teal_device_sm_count()
buf = torch.zeros(nsm * 16 + nsm * 64, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for shp in a.shapes.split(","):
    Z, N = (int(v) for v in shp.split("x"))
    W = (torch.randn(Z, N, device="cuda", dtype=torch.float16) * 0.02).T
    x = torch.randn(1, 1, Z, device="cuda", dtype=torch.float16)
    thr = float(x.abs().float().flatten().kthvalue(max(1, int(Z * a.sparsity))).values) if a.sparsity > 0 else -1.0
    for cfg in (int(c) for c in a.configs.split(",")):
        ops.set_config(cfg)
        ops.splitk_sparse_gemv(x, W, thr, 0)
        for rep in range(2):
            flush.zero_()
            buf.zero_()
            lib.teal_set_debug_timestamps(buf.data_ptr())
            ops.splitk_sparse_gemv(x, W, thr, 0)
            torch.cuda.synchronize()
            lib.teal_set_debug_timestamps(None)
        allb = buf.cpu().numpy()
        t = allb[: nsm * 16].reshape(nsm, 16)
        rowt = allb[nsm * 16:].reshape(nsm, 2, 32)
        for cta in (0, 73, 147):
            for wi, wn in ((0, "w0"), (1, "w15")):
                r = rowt[cta, wi]
                r = r[r > 0] - t[cta, 0]
                print(f"   cta {cta} {wn} row-done cycles: " + " ".join(str(int(v)) for v in r))
        rel = t[:, :13] - t[:, 0:1]
        g0 = t[:, 13].min()
        print(f"== {shp} cfg {cfg}: kernel span {(t[:,14].max()-g0)} ns; CTA start skew max {(t[:,13]-g0).max()} ns; "
              f"CTA end-start median {(np.median(t[:,14]-t[:,13]))} ns; units/CTA {t[0,15] >> 8}; "
              f"pieces/CTA hist {np.bincount((t[:,15] & 255).astype(int)).tolist()}")
        dur = t[:, 14] - t[:, 13]
        order = np.argsort(dur)
        for tag, ids in (("fastest", order[:3]), ("slowest", order[-4:])):
            for i in ids:
                print(f"   {tag} cta {i}: dur {dur[i]} ns np {t[i,15]&255} rel " + " ".join(str(int(v)) for v in rel[i, 1:13]))
        med = np.median(rel, axis=0)
        mx = rel.max(axis=0)
        print("   phase        median_cyc   max_cyc")
        for i, n in enumerate(NAMES):
            print(f"   {n:12s} {med[i]:10.0f} {mx[i]:10.0f}")
ops.set_config(0)
